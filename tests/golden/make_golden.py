"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference from /root/reference.

Run once in the build container (the reference tree does not exist on the GPU box):
    python tests/golden/make_golden.py
Every fixture stores seeds/inputs and the reference's outputs; tests regenerate the inputs from the same
functions in tests/synth.py and compare (a) the oracle restatements and (b) the CUDA path against them.
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle import pmf_oracle as po  # noqa: E402
from oracle.ref_loader import REFERENCE_ROOT, load_reference, load_reference_file  # noqa: E402
from tests import synth  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote %-28s %7.1f KB" % (name, os.path.getsize(path) / 1024))


def gen_knn(ref):
    for case in synth.KNN_CASES:
        inp = synth.knn_inputs(case)
        knn = ref.postproc.KNN(dict(knn=case["knn"], search=case["search"], sigma=case["sigma"], cutoff=case["cutoff"]),
                               case["nclasses"])
        out = knn(torch.from_numpy(inp["proj_range"]), torch.from_numpy(inp["unproj_range"]),
                  torch.from_numpy(inp["proj_argmax"]), torch.from_numpy(inp["px"]), torch.from_numpy(inp["py"]))
        save("knn_%s.npz" % case["name"], out=out.numpy().astype(np.int64))


def gen_project():
    # The reference's two functions, run unmodified: SemanticKitti.mapLidar2Camera (unbound, with a stub self)
    # and PerspectiveViewLoader.__getitem__ (with a stub dataset object).
    parser = load_reference_file("pc_processor/dataset/semantic_kitti/parser.py", "ref_parser")
    for name in ("pc_processor", "pc_processor.dataset", "pc_processor.dataset.preprocess"):
        if name not in sys.modules:
            m = types.ModuleType(name)
            m.__path__ = [os.path.join(REFERENCE_ROOT, *name.split("."))]
            sys.modules[name] = m
    loader_mod = load_reference_file("pc_processor/dataset/perspective_view_loader.py", "ref_pv_loader")
    for name in ("pc_processor", "pc_processor.dataset", "pc_processor.dataset.preprocess",
                 "pc_processor.dataset.preprocess.augmentor"):
        sys.modules.pop(name, None)

    for case in synth.PROJECT_CASES:
        inp = synth.project_inputs(case)
        H, W = case["H"], case["W"]

        class StubDataset:
            has_image = True
            proj_matrix = {0: inp["proj_matrix"]}

            def loadDataByIndex(self, index):
                return inp["pointcloud"], inp["labels"], None

            def loadImage(self, index):
                from PIL import Image
                return Image.fromarray(np.zeros((H, W, 3), np.uint8))

            def parsePathInfoByIndex(self, index):
                return 0, 0

            def labelMapping(self, label):
                return label

            mapLidar2Camera = parser.SemanticKitti.mapLidar2Camera

            def __len__(self):
                return 1

        cfg = {"augmentation": {}, "sensor": {"proj_h": H, "proj_w": W, "proj_ht": H, "proj_wt": W, "h_pad": 0, "w_pad": 0}}
        ld = loader_mod.PerspectiveViewLoader(StubDataset(), cfg, is_train=False, return_uproj=True)
        feat, mask, label, x_idx, y_idx, depth = ld[0]
        save("project_%s.npz" % case["name"], feat=feat[:5].numpy(), mask=mask.numpy(), label=label.numpy(),
             rows=x_idx.numpy(), cols=y_idx.numpy(), depth=depth.numpy())


def gen_fusion(ref):
    from pc_processor_ref_alias import pmf_net  # registered by main()
    for case in synth.FUSION_CASES:
        blk = pmf_net.ResidualBasedFusionBlock(case["pcd_c"], case["img_c"])
        shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
        sd = po.synth_state_dict(shapes, seed=case["seed"])
        blk.load_state_dict(sd)
        pcd, img = synth.fusion_inputs(case)
        blk.eval()
        with torch.no_grad():
            out_eval = blk(pcd, img)
        blk.train()
        out_train = blk(pcd, img)
        save("fusion_%s.npz" % case["name"], out_eval=out_eval.numpy(), out_train=out_train.detach().numpy())


def gen_pmf(ref):
    for case in synth.PMF_CASES:
        torch.manual_seed(1)
        m = ref.models.PMFNet(pcd_channels=5, img_channels=3, nclasses=case["nclasses"], base_channels=32,
                              imagenet_pretrained=False, image_backbone=case["backbone"])
        shapes = po.pmf_param_shapes(case["nclasses"], 32, case["backbone"])
        assert list(shapes.keys()) == list(m.state_dict().keys())
        sd = po.synth_state_dict(shapes, seed=case["seed"])
        m.load_state_dict(sd, strict=True)
        pcd, img = synth.pmf_inputs(case)
        m.eval()
        with torch.no_grad():
            lid_e, cam_e = m(pcd, img)
        # train mode: batch-stat BN, dropout disabled (modules in eval) so no RNG is involved
        m.train()
        for mod in m.modules():
            if isinstance(mod, torch.nn.Dropout2d):
                mod.eval()
        lid_t, cam_t = m(pcd, img)
        wl, wc = synth.pmf_loss_weights(case)
        loss = (lid_t * wl).sum() + (cam_t * wc).sum()
        loss.backward()
        gnorm = np.array([float(p.grad.double().norm()) if p.grad is not None else -1.0 for _, p in m.named_parameters()])
        gnames = np.array([n for n, _ in m.named_parameters()])
        picks = {}
        for n, p in m.named_parameters():
            if n in synth.PMF_GRAD_PICKS:
                picks["grad__" + n] = p.grad.numpy()
        sd_after = m.state_dict()
        stats = {("stat__" + k): sd_after[k].numpy() for k in synth.PMF_STAT_PICKS}
        save("pmf_%s.npz" % case["name"], lidar_eval=lid_e.numpy(), camera_eval=cam_e.numpy(),
             lidar_train=lid_t.detach().numpy(), camera_train=cam_t.detach().numpy(), loss=np.float64(loss.item()),
             grad_norms=gnorm, grad_names=gnames, **picks, **stats)


def gen_epmf(ref):
    from oracle import epmf_oracle as eo
    for case in synth.EPMF_CASES:
        torch.manual_seed(1)
        m = ref.models.EPMFNet(pcd_channels=5, img_channels=3, nclasses=case["nclasses"], base_channels=32,
                               imagenet_pretrained=False, image_backbone=case["backbone"])
        shapes = eo.epmf_param_shapes(case["nclasses"], 32, case["backbone"])
        assert list(shapes.keys()) == list(m.state_dict().keys())
        m.load_state_dict(po.synth_state_dict(shapes, seed=case["seed"]), strict=True)
        pcd, img = synth.epmf_inputs(case)
        m.eval()
        with torch.no_grad():
            lid, cam = m(pcd, img)
        save("epmf_%s.npz" % case["name"], lidar_eval=lid.numpy(), camera_eval=cam.numpy())


def main():
    ref = load_reference()
    import importlib
    alias = types.ModuleType("pc_processor_ref_alias")
    alias.pmf_net = importlib.import_module("ref_pc_processor.models.pmf_net")
    sys.modules["pc_processor_ref_alias"] = alias
    gen_knn(ref)
    gen_project()
    gen_fusion(ref)
    gen_pmf(ref)
    gen_epmf(ref)


if __name__ == "__main__":
    main()

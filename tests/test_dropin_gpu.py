"""Drop-in acceptance (-m gpu): the reference's BYTE-IDENTICAL ``tasks/pmf/main.py`` + ``option.py`` +
``trainer.py`` run end to end over our ``pc_processor`` shim on a synthetic SemanticKITTI-layout tree, under
``python -m torch.distributed.run`` exactly as ``tasks/pmf/run.sh`` launches them:

  Experiment -> pc_processor.utils.init_distributed_mode -> pc_processor.models.PMFNet (pmf_b200, libpmf_b200.so)
  -> Trainer: pc_processor.dataset.PerspectiveViewLoader (projection + scatter on the device, pmfb_project_scatter),
     the reference's Lovasz / focal / KL loss block, AdamW + SGD, pc_processor.layers.sync_bn.replaceBN, DDP, IOUEval
  -> one training iteration + one validation iteration (is_debug) -> checkpoint.

The reference tree is the staged copy baseline/_ref (tools/stage_reference.py; /root/reference does not exist on the GPU
box); everything is symlinked into a scratch directory (tests/merged_tree.py), nothing is modified.  The YAML is the
reference's own config with the data paths / sizes of the synthetic tree, ``n_threads: 0`` (the GPU loader runs in the main
process) and ``gpu: "0,1"``: ``Option`` counts the comma-separated entries, so the multi-GPU branch (replaceBN + DDP +
DistributedSampler + all-reduced IOUEval, trainer.py:33-39) runs with two ranks on a two-GPU box and with one rank on a
one-GPU box (CUDA ignores the visible-device entries after the first invalid index).
"""
import os
import re
import socket
import subprocess
import sys

import pytest
import torch
import yaml

from tests import merged_tree

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))

pytestmark = pytest.mark.gpu

H, W = 96, 160  # synthetic camera frames (the BASELINE 480x640 frame is exercised by bench.py; this test is about plumbing)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _sha(path):
    import hashlib
    return hashlib.sha256(open(path, "rb").read()).hexdigest()


@pytest.fixture(scope="module")
def tree(tmp_path_factory):
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    ref = merged_tree.reference_root()
    if ref is None:
        pytest.skip("no staged reference tree (run tools/stage_reference.py where /root/reference exists)")
    import make_synthetic_kitti as mk
    tmp = tmp_path_factory.mktemp("dropin")
    top = merged_tree.build(tmp, ref)
    data = os.path.join(str(tmp), "sequences")
    os.makedirs(data)
    for seq in range(11):  # trainer.py:103-124: sequences 0-7, 9, 10 train, 8 validation
        mk.write_sequence(data, seq, 2, H, W, rows=16, cols=1024, seed=3)
    return dict(top=top, data=data, ref=ref, tmp=str(tmp))


def _config(tree, nproc):
    cfg = yaml.safe_load(open(os.path.join(tree["ref"], "tasks", "pmf", "config_server_kitti.yaml")))
    cfg.update(save_path=os.path.join(tree["tmp"], "experiments") + os.sep, gpu="0,1",
               n_threads=0, is_debug=True, n_epochs=1, batch_size=[2, 2], data_root=tree["data"], imagenet_pretrained=False,
               experiment_id="dropin")
    # train crop 64x128 (+ pad 8) -> 80x144, validation centre crop = the whole 96x160 frame (+ pad 8) -> 112x176
    cfg["sensor"].update(proj_h=H + 16, proj_w=W + 16, proj_ht=80, proj_wt=144, h_pad=8, w_pad=8)
    path = os.path.join(tree["tmp"], "config_dropin.yaml")
    yaml.safe_dump(cfg, open(path, "w"))
    return path


def test_unchanged_main_py_trains_and_validates_on_our_kernels(tree):
    # the task scripts really are the reference's bytes
    for f in ("main.py", "option.py", "trainer.py"):
        assert _sha(os.path.join(tree["top"], "tasks", "pmf", f)) == _sha(os.path.join(tree["ref"], "tasks", "pmf", f))
        assert os.path.islink(os.path.join(tree["top"], "tasks", "pmf", f))
    nproc = 2 if torch.cuda.device_count() >= 2 else 1
    cfg = _config(tree, nproc)
    env = dict(os.environ)
    env["PYTHONPATH"] = os.path.join(tree["top"], "stubs") + os.pathsep + env.get("PYTHONPATH", "")
    env["PMFB_TRACE_LOADS"] = os.path.join(tree["tmp"], "loads.txt")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(nproc), "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "main.py", cfg]
    r = subprocess.run(cmd, cwd=os.path.join(tree["top"], "tasks", "pmf"), env=env, capture_output=True, text=True, timeout=900)
    out = r.stdout + "\n" + r.stderr
    try:  # keep the run's log next to the other GPU-box artefacts
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        open(os.path.join(ROOT, "gpurun_out", "dropin_main.log"), "w").write(out)
    except OSError:
        pass
    assert r.returncode == 0, out[-6000:]
    assert "===init env success===" in out
    # one training and one validation iteration were logged with finite numbers (trainer.py:415-426, 519-521)
    for mode in ("Train", "Validation"):
        m = re.search(r">>> %s Loss ([0-9.naninf-]+) Acc ([0-9.naninf-]+) IOU ([0-9.naninf-]+)" % mode, out)
        assert m, out[-4000:]
        loss, acc, iou = (float(x) for x in m.groups())
        assert loss == loss and 0.0 < loss < 100.0 and 0.0 <= acc <= 1.0 and 0.0 <= iou <= 1.0, (mode, m.groups())
    # the model that ran was ours, on the native library, with the device loader
    assert "libpmf_b200.so" in open(env["PMFB_TRACE_LOADS"]).read()
    # checkpoint written by main.py:104-113 loads strictly into a fresh module (same 654 keys)
    exp = os.path.join(tree["tmp"], "experiments")
    ckpts = [os.path.join(dp, f) for dp, _dn, fn in os.walk(exp) for f in fn if f == "checkpoint.pth"]
    assert len(ckpts) == 1, ckpts
    ck = torch.load(ckpts[0], map_location="cpu")
    assert set(ck) == {"model", "optimizer", "aux_optimizer", "epoch"} and ck["epoch"] == 0
    import pmf_b200
    fresh = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34")
    fresh.load_state_dict(ck["model"], strict=True)
    assert len(ck["model"]) == 654
    # the BN running statistics moved (train mode ran through the swapped-in SynchronizedBatchNorm2d containers)
    assert int(ck["model"]["lidar_stream.downCntx.bn1.num_batches_tracked"]) == 1
    assert float(ck["model"]["lidar_stream.downCntx.bn1.running_mean"].abs().sum()) > 0


def test_device_loader_equals_reference_loader(tree):
    """pmf_b200.loader.PerspectiveViewLoader against the reference's numpy loader on the same synthetic frames
    (validation mode: no random augmentation): identical (8,H,W) features, mask and labels, and the same un-projection
    tuple with return_uproj=True."""
    code = r'''
import os, sys, numpy as np, torch, yaml
sys.path.insert(0, %(stubs)r); os.chdir(%(cwd)r); sys.path.insert(0, "../../")
import pc_processor, importlib.util
spec = importlib.util.spec_from_file_location("ref_pvl", %(ref_loader)r)
ref_pvl = importlib.util.module_from_spec(spec); spec.loader.exec_module(ref_pvl)
cfg = yaml.safe_load(open(%(cfg)r))
ds = pc_processor.dataset.semantic_kitti.SemanticKitti(root=cfg["data_root"], sequences=[8],
        config_path="../../pc_processor/dataset/semantic_kitti/semantic-kitti.yaml")
assert pc_processor.dataset.PerspectiveViewLoader.__module__ == "pmf_b200.loader"
for kw in (dict(is_train=False, use_padding=True), dict(is_train=False, return_uproj=True)):
    ours = pc_processor.dataset.PerspectiveViewLoader(dataset=ds, config=cfg, **kw)
    theirs = ref_pvl.PerspectiveViewLoader(dataset=ds, config=cfg, **kw)
    assert len(ours) == len(theirs) == 2
    for i in range(len(ours)):
        a, b = ours[i], theirs[i]
        assert len(a) == len(b)
        for x, y in zip(a, b):
            assert x.is_cuda and tuple(x.shape) == tuple(y.shape), (x.shape, y.shape)
            xc = x.cpu().to(y.dtype)
            if not torch.equal(xc, y):
                d = (xc != y)
                per = d.reshape(d.shape[0], -1).sum(1).tolist() if d.dim() == 3 else int(d.sum())
                idx = d.nonzero()[:5].tolist()
                raise AssertionError("loader mismatch %%s: per-channel differing elements %%s, first at %%s: ours %%s theirs %%s" %% (
                    kw, per, idx, [float(xc[tuple(i)]) for i in idx], [float(y[tuple(i)]) for i in idx]))
print("loader ok")
''' % dict(stubs=os.path.join(tree["top"], "stubs"), cwd=os.path.join(tree["top"], "tasks", "pmf"), cfg=_config(tree, 1),
           ref_loader=os.path.join(tree["ref"], "pc_processor", "dataset", "perspective_view_loader.py"))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600)
    try:
        open(os.path.join(ROOT, "gpurun_out", "dropin_loader.log"), "w").write(r.stdout + "\n" + r.stderr)
    except OSError:
        pass
    assert r.returncode == 0 and "loader ok" in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])

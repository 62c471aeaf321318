"""The oracle restatements must reproduce the reference's outputs: the committed fixtures everywhere, and
the live reference whenever /root/reference is present (build container)."""
import os

import numpy as np
import pytest
import torch

from oracle import knn_oracle, pmf_oracle as po, project_oracle
from oracle.ref_loader import reference_available
from tests import synth

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    return np.load(os.path.join(GOLDEN, name))


@pytest.mark.parametrize("case", synth.KNN_CASES, ids=lambda c: c["name"])
def test_knn_oracle_matches_golden(case):
    inp = synth.knn_inputs(case)
    out, tie_free = knn_oracle.knn_vote(inp["proj_range"], inp["unproj_range"], inp["proj_argmax"], inp["px"], inp["py"],
                                        case["knn"], case["search"], case["sigma"], case["cutoff"], case["nclasses"],
                                        return_aux=True)
    ref = load("knn_%s.npz" % case["name"])["out"]
    assert out.shape == ref.shape and out.dtype == np.int64
    # bit-exact wherever the reference's unordered topk has a unique answer
    assert np.array_equal(out[tie_free], ref[tie_free])
    if case["kind"] != "ties":
        assert tie_free.mean() > 0.99
        assert (out != ref).mean() < 0.002


def test_knn_even_kernel_raises():
    inp = synth.knn_inputs(synth.KNN_CASES[0])
    with pytest.raises(ValueError, match="odd"):
        knn_oracle.knn_vote(inp["proj_range"], inp["unproj_range"], inp["proj_argmax"], inp["px"], inp["py"], 5, 4, 1.0, 1.0, 20)


@pytest.mark.parametrize("case", synth.PROJECT_CASES, ids=lambda c: c["name"])
def test_project_oracle_matches_golden(case):
    inp = synth.project_inputs(case)
    got = project_oracle.project_scatter(inp["proj_matrix"], inp["pointcloud"], inp["labels"], case["H"], case["W"])
    ref = load("project_%s.npz" % case["name"])
    assert np.array_equal(got["rows"], ref["rows"]) and np.array_equal(got["cols"], ref["cols"])
    assert np.array_equal(got["depth"], ref["feat"][0])
    assert np.array_equal(np.moveaxis(got["xyzi"], -1, 0), ref["feat"][1:5])
    assert np.array_equal(got["mask"].astype(np.float32), ref["mask"])
    assert np.array_equal(got["label"].astype(np.float32), ref["label"])
    assert np.array_equal(got["point_depth"], ref["depth"])
    assert got["mask"].sum() > 100


@pytest.mark.parametrize("case", synth.FUSION_CASES, ids=lambda c: c["name"])
def test_fusion_oracle_matches_golden(case):
    shapes = {}
    pc, ic = case["pcd_c"], case["img_c"]
    for name, cin in (("fuse_conv.0", pc + ic), ("attention.0", pc), ("attention.3", pc)):
        shapes[name + ".weight"] = (pc, cin, 3, 3)
        shapes[name + ".bias"] = (pc,)
    for name in ("fuse_conv.2", "attention.1", "attention.4"):
        for k, s in (("weight", (pc,)), ("bias", (pc,)), ("running_mean", (pc,)), ("running_var", (pc,)), ("num_batches_tracked", ())):
            shapes[name + "." + k] = s
    sd = po.synth_state_dict(shapes, seed=case["seed"])
    sd = {"blk." + k: v for k, v in sd.items()}
    pcd, img = synth.fusion_inputs(case)
    ref = load("fusion_%s.npz" % case["name"])
    out = po.fusion_block(po.Ctx(sd), pcd, img, "blk")
    assert torch.allclose(out, torch.from_numpy(ref["out_eval"]), atol=1e-6, rtol=1e-6)
    out = po.fusion_block(po.Ctx(sd, train=True), pcd, img, "blk")
    assert torch.allclose(out, torch.from_numpy(ref["out_train"]), atol=1e-5, rtol=1e-5)


@pytest.mark.parametrize("case", synth.PMF_CASES, ids=lambda c: c["name"])
def test_pmf_oracle_matches_golden(case):
    shapes = po.pmf_param_shapes(case["nclasses"], 32, case["backbone"])
    sd = po.synth_state_dict(shapes, seed=case["seed"])
    pcd, img = synth.pmf_inputs(case)
    ref = load("pmf_%s.npz" % case["name"])
    with torch.no_grad():
        lid, cam = po.pmf_forward(sd, pcd, img, case["backbone"])
    assert (lid - torch.from_numpy(ref["lidar_eval"])).abs().max() < 1e-6
    assert (cam - torch.from_numpy(ref["camera_eval"])).abs().max() < 1e-6
    params = {k: v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v for k, v in sd.items()}
    lid, cam, ctx = po.pmf_forward(params, pcd, img, case["backbone"], train=True, return_ctx=True)
    assert (lid - torch.from_numpy(ref["lidar_train"])).abs().max() < 1e-5
    assert (cam - torch.from_numpy(ref["camera_train"])).abs().max() < 1e-5
    wl, wc = synth.pmf_loss_weights(case)
    loss = (lid * wl).sum() + (cam * wc).sum()
    assert abs(loss.item() - float(ref["loss"])) < 1e-3 * max(1.0, abs(float(ref["loss"])))
    loss.backward()
    names = [str(n) for n in ref["grad_names"]]
    for n, gn in zip(names, ref["grad_norms"]):
        g = params[n].grad
        assert g is not None
        assert abs(float(g.double().norm()) - gn) <= 2e-3 * max(gn, 1e-6) + 1e-7, n
    for n in synth.PMF_GRAD_PICKS:
        g_ref = torch.from_numpy(ref["grad__" + n])
        assert (params[n].grad - g_ref).abs().max() <= 2e-3 * g_ref.abs().max() + 1e-7, n
    for k in synth.PMF_STAT_PICKS:
        assert torch.allclose(ctx.new_stats[k], torch.from_numpy(ref["stat__" + k]), atol=1e-5, rtol=1e-5), k


@pytest.mark.parametrize("case", synth.EPMF_CASES, ids=lambda c: c["name"])
def test_epmf_oracle_matches_golden(case):
    from oracle import epmf_oracle as eo
    sd = po.synth_state_dict(eo.epmf_param_shapes(case["nclasses"], 32, case["backbone"]), seed=case["seed"])
    pcd, img = synth.epmf_inputs(case)
    ref = load("epmf_%s.npz" % case["name"])
    with torch.no_grad():
        lid, cam = eo.epmf_forward(sd, pcd, img, case["backbone"])
    assert (lid - torch.from_numpy(ref["lidar_eval"])).abs().max() < 1e-6
    assert (cam - torch.from_numpy(ref["camera_eval"])).abs().max() < 1e-6


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_epmf_oracle_equals_live_reference():
    """EPMFNet: same state_dict inventory (689 keys, reference order), same default init under the same seed for our
    module tree, and a bit-identical eval forward of the oracle on the reference's own initialisation."""
    from oracle import epmf_oracle as eo
    from oracle.ref_loader import load_reference
    from pmf_b200 import modules as M
    ref = load_reference()
    torch.manual_seed(1)
    m = ref.models.EPMFNet(5, 3, 20, 32, False, "resnet34")
    shapes = eo.epmf_param_shapes(20, 32, "resnet34")
    sd = m.state_dict()
    assert list(shapes.keys()) == list(sd.keys())
    assert all(tuple(v.shape) == tuple(shapes[k]) for k, v in sd.items())
    torch.manual_seed(1)
    ours = M.EPMFNet(5, 3, 20, 32, False, "resnet34").state_dict()
    assert list(ours.keys()) == list(sd.keys()) and all(torch.equal(ours[k], sd[k]) for k in sd)
    feat, _, _ = synth.frame_tensor(1, 32, 64, seed=6, density=0.2)
    m.eval()
    with torch.no_grad():
        a, b = m(feat[:, :5], feat[:, 5:8])
        a2, b2 = eo.epmf_forward(sd, feat[:, :5], feat[:, 5:8])
    assert torch.equal(a, a2) and torch.equal(b, b2)


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_oracle_equals_live_reference():
    from oracle.ref_loader import load_reference
    ref = load_reference()
    torch.manual_seed(1)
    m = ref.models.PMFNet(5, 3, 20, 32, False, "resnet34")
    shapes = po.pmf_param_shapes(20, 32, "resnet34")
    assert list(shapes.keys()) == list(m.state_dict().keys())
    assert all(tuple(v.shape) == tuple(shapes[k]) for k, v in m.state_dict().items())
    sd = m.state_dict()  # the reference's own default init
    feat, _, _ = synth.frame_tensor(1, 16, 32, seed=5)
    m.eval()
    with torch.no_grad():
        a, b = m(feat[:, :5], feat[:, 5:8])
        a2, b2 = po.pmf_forward(sd, feat[:, :5], feat[:, 5:8])
    assert torch.equal(a, a2) and torch.equal(b, b2)
    # KNN: oracle vs the reference module on a fresh random case
    case = dict(synth.KNN_CASES[0], seed=99, P=500)
    inp = synth.knn_inputs(case)
    knn = ref.postproc.KNN(dict(knn=5, search=5, sigma=1.0, cutoff=1.0), 20)
    r = knn(*[torch.from_numpy(inp[k]) for k in ("proj_range", "unproj_range", "proj_argmax", "px", "py")]).numpy()
    o, tf = knn_oracle.knn_vote(inp["proj_range"], inp["unproj_range"], inp["proj_argmax"], inp["px"], inp["py"], 5, 5, 1.0, 1.0, 20, return_aux=True)
    assert np.array_equal(o[tf], r[tf])


@pytest.mark.skipif(not reference_available(), reason="reference tree not present (GPU box)")
def test_loss_oracle_equals_reference_modules():
    """the trainer's loss block (tasks/pmf/trainer.py:188-252, 305-332) composed from the reference's own
    FocalSoftmaxLoss / Lovasz_softmax / nn.KLDivLoss exactly as the trainer composes them, against oracle/loss_oracle.py:
    value and gradients w.r.t. both probability maps."""
    import math
    from oracle import loss_oracle as lo
    from oracle.ref_loader import load_reference
    ref = load_reference()
    torch.manual_seed(3)
    C, tau, gamma_w, lam = 20, 0.7, 0.5, 1.0
    label = torch.randint(0, C, (2, 24, 40))
    label[0, :6] = 0  # unlabelled rows
    cls_freq = np.linspace(0.01, 0.2, C)
    alpha_np = np.log(1 + 1.0 / (cls_freq + 1e-3))
    alpha_np = alpha_np / alpha_np.max()
    alpha_np[0] = 0
    focal = ref.loss.FocalSoftmaxLoss(C, gamma=2, alpha=alpha_np, softmax=False)
    lovasz = ref.loss.Lovasz_softmax(ignore=0)
    kl = torch.nn.KLDivLoss(reduction="none")

    def reference_total(lidar_pred, camera_pred):
        mask = label.gt(0)
        lp_log = torch.log(lidar_pred.clamp(min=1e-8))
        pcd_entropy = -(lidar_pred * lp_log).sum(1) / math.log(C)
        cp_log = torch.log(camera_pred.clamp(min=1e-8))
        img_entropy = -(camera_pred * cp_log).sum(1) / math.log(C)
        pcd_conf, img_conf = 1 - pcd_entropy, 1 - img_entropy
        imp = pcd_conf - img_conf
        pcd_w = imp.gt(0).float() * imp.abs() * pcd_conf.ge(tau).float()
        img_w = imp.lt(0).float() * imp.abs() * img_conf.ge(tau).float()
        per = (kl(lp_log, camera_pred) * img_w.unsqueeze(1)).mean() + (kl(cp_log, lidar_pred) * pcd_w.unsqueeze(1)).mean()
        return (focal(lidar_pred, label, mask=mask) + lam * lovasz(lidar_pred, label) +
                focal(camera_pred, label, mask=mask) + lam * lovasz(camera_pred, label) + gamma_w * per)

    # peaked predictions so that a good share of the pixels passes the tau = 0.7 confidence threshold
    la = (torch.randn(2, C, 24, 40) * 4).requires_grad_(True)
    lb = (torch.randn(2, C, 24, 40) * 4).requires_grad_(True)
    r = reference_total(torch.softmax(la, 1), torch.softmax(lb, 1))
    ga, gb = torch.autograd.grad(r, (la, lb))
    la2, lb2 = la.detach().clone().requires_grad_(True), lb.detach().clone().requires_grad_(True)
    o = lo.total_loss(torch.softmax(la2, 1), torch.softmax(lb2, 1), label, torch.from_numpy(alpha_np).float(), C, lam, gamma_w, tau)
    ga2, gb2 = torch.autograd.grad(o, (la2, lb2))
    assert abs(float(o) - float(r)) < 1e-5 * max(1.0, abs(float(r)))
    assert torch.allclose(ga2, ga, rtol=1e-4, atol=1e-7) and torch.allclose(gb2, gb, rtol=1e-4, atol=1e-7)
    per, pcd_w, img_w = lo.perception_aware_loss(torch.softmax(la, 1).detach(), torch.softmax(lb, 1).detach(), C, tau)
    assert float(pcd_w.gt(0).float().mean()) > 0.02 and float(img_w.gt(0).float().mean()) > 0.02  # the KL terms are exercised
    assert torch.allclose(lo.focal_alpha(cls_freq), torch.from_numpy(alpha_np).float(), rtol=1e-6, atol=1e-7)

"""Host-logic tests (no GPU): pmf_b200.engine / pmf_b200.net driven through a numpy model of the C-ABI
(tests/cabi_mock.py) and compared with the oracle.  This checks everything ABOVE the boundary — graph wiring,
implicit-GEMM tap generation (dilation, 2x2-dilated, stride-2 parity layout, 7x7 stem unrolling), the stride-2
dgrad decomposition, gradient-slice bookkeeping of the concat buffers and the BatchNorm backward plumbing.
The kernels themselves are checked by the -m gpu tests."""
import numpy as np
import pytest
import torch

from oracle import pmf_oracle as po
from pmf_b200 import modules as M
from tests import cabi_mock, synth


def _maxrel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def _load(model, sd):
    model.load_state_dict(sd, strict=True)
    return model


@pytest.fixture
def mock(monkeypatch):
    """tf32-rounding model of the library (what the B200 computes)."""
    cabi_mock.install(monkeypatch)


@pytest.fixture
def mock_exact(monkeypatch):
    """Exact-fp32 model of the library: isolates wiring / indexing / formula errors from tf32 noise."""
    cabi_mock.install(monkeypatch, exact=True)


def _grad_err(g, r, scale):
    """max |g - r| relative to the larger of max|r| and `scale` (bias gradients in front of a BatchNorm are
    analytically zero: compare them on the scale of the layer's weight gradient instead of their own)."""
    return float((g - r).abs().max() / max(float(r.abs().max()), scale))


def _l2_err(g, r, scale):
    """Relative L2 error.  Whole-network gradients are compared in L2 because a single activation sitting within
    fp32 noise of a LeakyReLU/ReLU kink may take the other branch in the two implementations, which changes ONE
    element of one gradient tensor by O(1) (and everything upstream by a fraction of a percent) without either
    side being wrong; block-level tests above use the max norm."""
    return float((g - r).double().norm() / max(float(r.double().norm()), scale))


def test_fusion_block_matches_oracle(mock_exact):
    case = synth.FUSION_CASES[0]
    blk = M.ResidualBasedFusionBlock(case["pcd_c"], case["img_c"])
    shapes = {k: tuple(v.shape) for k, v in blk.state_dict().items()}
    sd = po.synth_state_dict(shapes, seed=case["seed"])
    blk.load_state_dict(sd)
    pcd, img = synth.fusion_inputs(case)
    sdp = {"b." + k: v for k, v in sd.items()}
    # eval
    blk.eval()
    with torch.no_grad():
        out = M._FusionFn.apply(blk, False, pcd, img, *[p for _, p in blk.named_parameters()])
    ref = po.fusion_block(po.Ctx(sdp), pcd, img, "b")
    assert _maxrel(out, ref) < 1e-5
    # train + backward
    blk.train()
    pcd_g, img_g = pcd.clone().requires_grad_(True), img.clone().requires_grad_(True)
    out = M._FusionFn.apply(blk, True, pcd_g, img_g, *[p for _, p in blk.named_parameters()])
    rs = np.random.RandomState(0)
    dout = torch.from_numpy(rs.normal(0, 1, tuple(out.shape)).astype(np.float32))
    out.backward(dout)
    params = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in sdp.items()}
    pr, ir = pcd.clone().requires_grad_(True), img.clone().requires_grad_(True)
    ctx = po.Ctx(params, train=True)
    ref = po.fusion_block(ctx, pr, ir, "b")
    assert _maxrel(out.detach(), ref.detach()) < 1e-5
    ref.backward(dout)
    assert _maxrel(pcd_g.grad, pr.grad) < 1e-4
    assert _maxrel(img_g.grad, ir.grad) < 1e-4
    for n, p in blk.named_parameters():
        assert p.grad is not None, n
        scale = 1e-2 * float(params["b." + n.rsplit(".", 1)[0] + ".weight"].grad.abs().max())
        assert _grad_err(p.grad, params["b." + n].grad, scale) < 1e-3, n
    # running statistics were updated in place like nn.BatchNorm2d.train()
    for k, v in ctx.new_stats.items():
        assert torch.allclose(blk.state_dict()[k[2:]], v, atol=1e-5, rtol=1e-4), k
    assert int(blk.state_dict()["fuse_conv.2.num_batches_tracked"]) == 1


def _pmf_case(backbone, nclasses, B, H, W, seed):
    torch.manual_seed(0)
    m = M.PMFNet(5, 3, nclasses, 32, False, backbone)
    shapes = po.pmf_param_shapes(nclasses, 32, backbone)
    assert list(shapes.keys()) == list(m.state_dict().keys())
    sd = po.synth_state_dict(shapes, seed=seed)
    m.load_state_dict(sd, strict=True)
    feat, _, _ = synth.frame_tensor(B, H, W, seed=100 + seed)
    return m, sd, feat[:, 0:5], feat[:, 5:8]


def test_pmf_r34_eval_matches_oracle_and_golden(mock):
    case = synth.PMF_CASES[0]
    m, sd, pcd, img = _pmf_case(case["backbone"], case["nclasses"], case["B"], case["H"], case["W"], case["seed"])
    m.eval()
    with torch.no_grad():
        lid, cam = M._PMFFn.apply(m, False, pcd, img, *[p for _, p in m.named_parameters()])
        rl, rc = po.pmf_forward(sd, pcd, img, case["backbone"], tf32=True)
    # tf32 operand rounding on the He-scaled synthetic weights of this fixture is itself ~2e-3 of max p
    # (oracle-with-tf32-operands vs fp32 oracle: 1.8e-3); wiring errors would be O(1).
    assert _maxrel(lid, rl) < 5e-3 and _maxrel(cam, rc) < 5e-3
    gold = np.load("tests/golden/pmf_%s.npz" % case["name"])
    assert _maxrel(lid, torch.from_numpy(gold["lidar_eval"])) < 5e-3
    assert _maxrel(cam, torch.from_numpy(gold["camera_eval"])) < 5e-3


def test_pmf_r34_eval_exact_model_matches_golden(mock_exact):
    """With exact operands the executor reproduces the reference's fp32 outputs to accumulation-order noise."""
    case = synth.PMF_CASES[0]
    m, sd, pcd, img = _pmf_case(case["backbone"], case["nclasses"], case["B"], case["H"], case["W"], case["seed"])
    m.eval()
    with torch.no_grad():
        lid, cam = M._PMFFn.apply(m, False, pcd, img, *[p for _, p in m.named_parameters()])
    gold = np.load("tests/golden/pmf_%s.npz" % case["name"])
    assert _maxrel(lid, torch.from_numpy(gold["lidar_eval"])) < 2e-5
    assert _maxrel(cam, torch.from_numpy(gold["camera_eval"])) < 2e-5


def test_pmf_r34_train_backward_matches_oracle(mock_exact):
    case = synth.PMF_CASES[0]
    m, sd, pcd, img = _pmf_case(case["backbone"], case["nclasses"], case["B"], case["H"], case["W"], case["seed"])
    m.train()
    m._dropout_override = False
    lid, cam = M._PMFFn.apply(m, True, pcd, img, *[p for _, p in m.named_parameters()])
    wl, wc = synth.pmf_loss_weights(case)
    ((lid * wl).sum() + (cam * wc).sum()).backward()
    params = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in sd.items()}
    rl, rc, ctx = po.pmf_forward(params, pcd, img, case["backbone"], train=True, return_ctx=True)
    assert _maxrel(lid.detach(), rl.detach()) < 1e-4 and _maxrel(cam.detach(), rc.detach()) < 1e-4
    gold = np.load("tests/golden/pmf_%s.npz" % case["name"])
    assert _maxrel(lid.detach(), torch.from_numpy(gold["lidar_train"])) < 1e-4
    ((rl * wl).sum() + (rc * wc).sum()).backward()
    bad = []
    for n, p in m.named_parameters():
        assert p.grad is not None, n
        r = params[n].grad
        wname = n.rsplit(".", 1)[0] + ".weight"
        scale = 1e-2 * float(params[wname].grad.double().norm())
        err = _l2_err(p.grad, r, scale)
        if err > 2e-2:
            bad.append((n, err))
    assert not bad, bad[:10]
    for k, v in ctx.new_stats.items():
        assert torch.allclose(m.state_dict()[k], v, atol=1e-4, rtol=1e-3), k


def test_pmf_dropout_masks_match_oracle(mock_exact):
    """Explicit Dropout2d masks at every site: folded masks (pool / pixel-shuffle / concat copy / BN apply)."""
    m, sd, pcd, img = _pmf_case("resnet34", 20, 2, 32, 64, 3)
    m.train()
    rs = np.random.RandomState(5)
    sites = {"camera_stream_encoder.dropout.layer3": 256, "camera_stream_encoder.dropout.layer4": 512}
    for i, c in ((2, 128), (3, 256), (4, 256), (5, 256)):
        sites["lidar_stream.resBlock%d.dropout" % i] = c
    for i, (ca, cs, c) in ((1, (64, 256, 128)), (2, (32, 256, 128)), (3, (32, 128, 64))):
        sites["lidar_stream.upBlock%d.dropout1" % i] = ca
        sites["lidar_stream.upBlock%d.dropout2" % i] = ca + cs
        sites["lidar_stream.upBlock%d.dropout3" % i] = c
    masks = {k: torch.from_numpy(((rs.rand(2, c, 1, 1) < 0.8) / 0.8).astype(np.float32)) for k, c in sites.items()}
    m._dropout_override = {k: v.reshape(2, -1) for k, v in masks.items()}
    lid, cam = M._PMFFn.apply(m, True, pcd, img, *[p for _, p in m.named_parameters()])
    (lid[:, 3].sum() + cam[:, 5].sum()).backward()
    params = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in sd.items()}
    rl, rc = po.pmf_forward(params, pcd, img, "resnet34", train=True, dropout=masks)
    assert _maxrel(lid.detach(), rl.detach()) < 1e-4 and _maxrel(cam.detach(), rc.detach()) < 1e-4
    (rl[:, 3].sum() + rc[:, 5].sum()).backward()
    for n in ("lidar_stream.downCntx.conv2.weight", "camera_stream_encoder.layer3.0.conv1.weight",
              "lidar_stream.upBlock2.conv1.weight", "lidar_stream.resBlock3.conv5.bias"):
        r = params[n].grad
        assert _l2_err(dict(m.named_parameters())[n].grad, r, 1e-12) < 1e-1, n


def test_pmf_r50_17classes_eval(mock_exact):
    """Bottleneck encoder, 64-wide decoder, class count that is not a multiple of 4 (nuScenes: 17)."""
    m, sd, pcd, img = _pmf_case("resnet50", 17, 1, 16, 32, 7)
    m.eval()
    with torch.no_grad():
        lid, cam = M._PMFFn.apply(m, False, pcd, img, *[p for _, p in m.named_parameters()])
        rl, rc = po.pmf_forward(sd, pcd, img, "resnet50")
    assert lid.shape == (1, 17, 16, 32)
    assert _maxrel(lid, rl) < 2e-5 and _maxrel(cam, rc) < 2e-5


def test_invalid_input_size_asserts(mock):
    m, sd, pcd, img = _pmf_case("resnet34", 20, 1, 16, 16, 1)
    with pytest.raises(AssertionError, match="invalid input size"):
        M._PMFFn.apply(m, False, pcd[:, :, :, :12], img[:, :, :, :12], *[p for _, p in m.named_parameters()])


def test_modules_refuse_cpu_tensors():
    m = M.ResidualBasedFusionBlock(8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 8, 4, 4), torch.zeros(1, 8, 4, 4))


# ------------------------------------------------------------------------------------------------ EPMF (inference)
def _epmf_case(B, H, W, seed, density=0.35):
    from oracle import epmf_oracle as eo
    torch.manual_seed(seed)
    m = M.EPMFNet(5, 3, 20, 32, False, "resnet34")
    sd = po.synth_state_dict(eo.epmf_param_shapes(20, 32, "resnet34"), seed=seed)
    _load(m, sd)
    feat, _, _ = synth.frame_tensor(B, H, W, seed=seed + 100, density=density)
    return m, sd, feat[:, 0:5].contiguous(), feat[:, 5:8].contiguous()


@pytest.mark.parametrize("density", [0.35, 0.03])
def test_epmf_eval_matches_oracle(mock_exact, density):
    """EPMFNet eval forward (sparse context blocks with their mask dilation incl. the stride-2 one, fusion-before-
    ResBlock wiring at half resolution, extraUpSample, ASPP-fused camera decoder) against the oracle, exact arithmetic.
    density 0.03: most 3x3 windows are empty, so the dilated masks actually zero large regions."""
    from oracle import epmf_oracle as eo
    from pmf_b200 import net as G
    from pmf_b200.engine import Engine, WeightCache
    m, sd, pcd, img = _epmf_case(1, 32, 64, 11, density)
    m.eval()
    with torch.no_grad():
        E = Engine(G.ModuleParams(m), pcd.device, False, False, WeightCache(), dropout=False)
        lid, cam, _, _ = G.epmf_forward(E, pcd, img, "resnet34", 20)
        rl, rc = eo.epmf_forward(sd, pcd, img, "resnet34")
    assert lid.shape == (1, 20, 32, 64) and cam.shape == (1, 20, 32, 64)
    assert _maxrel(lid, rl) < 2e-5 and _maxrel(cam, rc) < 2e-5, (_maxrel(lid, rl), _maxrel(cam, rc))


def test_epmf_rejects_training_and_bad_sizes():
    m = M.EPMFNet(5, 3, 20, 32, False, "resnet34")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        m(torch.zeros(1, 5, 32, 32), torch.zeros(1, 3, 32, 32))


def test_precise_mode_3xtf32_recovers_fp32_on_the_truncating_model(mock):
    """PMFB_PRECISION=3xtf32 wiring: on the tf32 model of the library (the tensor core ignores the low 13 operand mantissa
    bits) the default mode deviates from the fp32 oracle by operand rounding (~1e-3 on this fixture) while the precise mode
    -- nothing rounded where it is produced, every conv operand split into [hi|hi|lo] x [hi|lo|hi] over 3x the channels,
    three wgrad launches -- matches it like exact arithmetic does, forward and gradients, stride-2 and stem included."""
    import pmf_b200
    case = synth.PMF_CASES[0]
    res = {}
    for mode in ("tf32", "3xtf32"):
        m, sd, pcd, img = _pmf_case(case["backbone"], case["nclasses"], case["B"], case["H"], case["W"], case["seed"])
        m.train()
        m._dropout_override = False
        with pmf_b200.precision(mode):
            lid, cam = M._PMFFn.apply(m, True, pcd, img, *[p for _, p in m.named_parameters()])
            wl, wc = synth.pmf_loss_weights(case)
            ((lid * wl).sum() + (cam * wc).sum()).backward()
        params = {k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
                  for k, v in sd.items()}
        rl, rc = po.pmf_forward(params, pcd, img, case["backbone"], train=True)
        ((rl * wl).sum() + (rc * wc).sum()).backward()
        errs = []
        for n, p in m.named_parameters():
            wname = n.rsplit(".", 1)[0] + ".weight"
            errs.append(_l2_err(p.grad, params[n].grad, 1e-2 * float(params[wname].grad.double().norm())))
        res[mode] = (max(_maxrel(lid.detach(), rl.detach()), _maxrel(cam.detach(), rc.detach())), float(np.median(errs)))
    # gradients: the same 2e-2 relative-L2 bar the exact-arithmetic wiring test uses (batch-statistics BN on random weights
    # amplifies even fp32 summation-order noise); the default mode sits at ~0.5 on this fixture
    assert res["3xtf32"][0] < 1e-4 and res["3xtf32"][1] < 2e-2, res
    assert res["tf32"][1] > 10 * res["3xtf32"][1], res
    assert res["tf32"][0] > 10 * res["3xtf32"][0], res  # the model really does lose the low operand bits in the default mode


def test_backward_order_with_a_forward_branch():
    """Engine._backward_order on a synthetic tape shaped like PMF's: camera closures [0, 14) = stem+layer1..layer4 (marks at
    3, 6, 9, 12) + decoder [12, 14), LiDAR closures [14, 30) with the four feature-map copies at 16, 20, 24, 28.  Every tape
    index runs exactly once, in an order that respects the hand-over rules documented in the method."""
    from pmf_b200 import engine as eng

    class Root:
        ready = object()

    roots = [Root() for _ in range(4)]
    tape = [lambda: None for _ in range(30)]
    W = [16, 20, 24, 28]
    for r, wi in zip(roots, W):
        def cp():
            return None
        cp.copy_src_root = r
        tape[wi] = cp

    class Stub:
        pass
    E = Stub()
    E.tape, E.use_branch = tape, True
    E._branch_span = [0, 14]
    E._marks = list(zip(roots, [3, 6, 9, 12]))
    E._waits = list(zip(roots, W))
    old = eng.BWD_BRANCH
    try:
        eng.BWD_BRANCH = True
        order = eng.Engine._backward_order(E)
        eng.BWD_BRANCH = False
        plain = eng.Engine._backward_order(E)
    finally:
        eng.BWD_BRANCH = old
    assert [e[1] for e in plain] == list(range(29, -1, -1)) and all(e[2] == "main" for e in plain)
    fns = [e for e in order if e[0] == "fn"]
    assert sorted(e[1] for e in fns) == list(range(30))
    pos = {e[1]: i for i, e in enumerate(order) if e[0] == "fn"}
    tag = {e[1]: e[2] for e in fns}
    assert all(tag[i] == "cam" for i in range(14)) and all(tag[i] == "main" for i in range(14, 30))
    # each chain keeps its own reversed order
    for chain in (range(14), range(14, 30)):
        idx = sorted(chain, key=lambda i: pos[i])
        assert idx == sorted(idx, reverse=True)
    assert pos[13] < pos[12] < min(pos[i] for i in range(12))          # decoder first on the camera chain
    wait_d = order.index(("wait", "main", "D"))
    assert order.index(("rec", "cam", "D")) < wait_d < pos[28]         # the LiDAR chain waits for it before its first copy
    for j, w in ((4, 24), (3, 20), (2, 16)):                             # layer j released after W[j-2]
        lo, hi = [3, 6, 9, 12][j - 2], [3, 6, 9, 12][j - 1]
        first = min(pos[i] for i in range(lo, hi))
        assert pos[w] < order.index(("rec", "main", "W%d" % w)) < order.index(("wait", "cam", "W%d" % w)) < first
    assert min(pos[i] for i in range(0, 3)) > max(pos[i] for i in range(3, 6))   # layer 1 + stem after layer 2
    # a broken hand-over record (copy closure not where expected) falls back to the plain order
    E._waits[1] = (roots[1], 21)
    eng.BWD_BRANCH = True
    try:
        assert eng.Engine._backward_order(E) == plain
    finally:
        eng.BWD_BRANCH = old

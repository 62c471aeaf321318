"""fp16 storage of the pre-BatchNorm activation ("f16" mode, pmfb_conv_desc.out_half / in_half / x_half): the BatchNorm
passes reading the fp16 buffer must give exactly what they give on the same values held in fp32, and the conv epilogue
must store the rounded fp32 result with the statistics of the stored values."""
import ctypes as C

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a B200")
    from pmf_b200 import _lib as L
    L.require_device()
    return torch.device("cuda:0")


def _null():
    from pmf_b200._lib import View
    return View()


@pytest.mark.parametrize("shape", [(2, 24, 40, 32), (3, 17, 23, 64), (1, 9, 11, 8), (2, 30, 40, 256)])
@pytest.mark.parametrize("variant", ["plain", "mask", "relu_z", "gate", "gacc"])
def test_bn_passes_fp16_input_equals_fp32_input(dev, shape, variant):
    from pmf_b200 import _lib as L
    from pmf_b200._lib import ACT_NONE, ACT_RELU, ACT_SIGMOID, Epilogue
    from pmf_b200.engine import _chan_view, _view
    n, h, w, c = shape
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(hash((shape, variant)) % 1000)
    a16 = (torch.randn(n, h, w, c, generator=g) * 2).half().to(dev)
    a32 = a16.float()
    dy = torch.randn(n, h, w, c, generator=g).to(dev)
    vec = (torch.rand(5 * c, generator=g) + 0.5).to(dev)
    alpha, beta, mean, invstd, gamma = (vec[i * c:(i + 1) * c].contiguous() for i in range(5))
    mask = (torch.rand(n, c, generator=g) > 0.3).float().mul(1.25).to(dev)
    f = torch.randn(n, h, w, c, generator=g).to(dev)
    shortcut = torch.randn(n, h, w, c, generator=g).to(dev)
    mulv = _chan_view(mask) if variant == "mask" else (_view(f) if variant == "gate" else _null())
    act_z = {"relu_z": ACT_RELU, "gate": ACT_SIGMOID}.get(variant, ACT_NONE)
    leaky_x = 1 if variant in ("plain", "mask", "gacc") else 0
    # ---- BN apply (pointwise16)
    outs = []
    for half in (0, 1):
        a = a16 if half else a32
        y = torch.empty(n, h, w, c, device=dev)
        y16 = torch.empty(n, h, w, c, device=dev, dtype=torch.float16)
        yb = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        e = Epilogue()
        e.alpha1, e.beta1, e.round_out = alpha.data_ptr(), beta.data_ptr(), 1
        if variant == "mask":
            e.mul = mulv
        if variant == "gacc":
            e.r1 = _view(shortcut)
        if variant == "relu_z":
            e.act = ACT_RELU
        if variant == "gate":
            e.act, e.mul, e.r2 = ACT_SIGMOID, _view(f), _view(shortcut)
        L.call("pmfb_pointwise16", C.byref(_view(a)), y.data_ptr(), c * h * w, c * w, c, n, h, w, c, C.byref(e),
               y16.data_ptr(), L.DT_F16, yb.data_ptr(), half, st)
        outs.append((y, y16, yb))
    for t0, t1 in zip(*outs):
        assert torch.equal(t0, t1)
    z = outs[0][0] if variant == "relu_z" else None
    # ---- backward reduce + apply
    res = []
    for half in (0, 1):
        a = a16 if half else a32
        red = torch.zeros(2 * c, device=dev, dtype=torch.float64)
        L.call("pmfb_bn_bwd_reduce16", C.byref(_view(dy)), C.byref(mulv), C.byref(_view(z)), act_z, C.byref(_view(a)),
               mean.data_ptr(), invstd.data_ptr(), alpha.data_ptr(), beta.data_ptr(), n, h, w, c, red.data_ptr(), half, st)
        dx = torch.empty(n, h, w, c, device=dev)
        dxb = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        gw, gb = torch.empty(c, device=dev), torch.empty(c, device=dev)
        cs = torch.zeros(c, device=dev, dtype=torch.float64)
        g_out = shortcut.clone() if variant == "gacc" else None
        L.call("pmfb_bn_bwd_apply16", C.byref(_view(dy)), C.byref(mulv), C.byref(_view(z)), act_z, C.byref(_view(a)),
               mean.data_ptr(), invstd.data_ptr(), alpha.data_ptr(), beta.data_ptr(), gamma.data_ptr(), red.data_ptr(), leaky_x,
               n, h, w, c, dx.data_ptr(), c * h * w, c * w, c, 1, gw.data_ptr(), gb.data_ptr(), cs.data_ptr(),
               None if g_out is None else g_out.data_ptr(), c * h * w, c * w, c, 1 if g_out is not None else 0,
               dxb.data_ptr(), half, st)
        res.append((red, dx, dxb, gw, gb, cs, g_out))
    torch.cuda.synchronize()
    r0, r1 = res
    # the element-wise results are the same arithmetic on the same values; the reductions differ only by summation order
    assert torch.allclose(r0[0], r1[0], rtol=1e-6, atol=1e-6 * n * h * w)
    assert torch.allclose(r0[1], r1[1], rtol=1e-5, atol=1e-5)
    assert torch.allclose(r0[5], r1[5], rtol=1e-6, atol=1e-6 * n * h * w)
    if variant == "gacc":
        assert torch.equal(r0[6], r1[6])
    # and against plain torch on the fp32 values
    gt = dy * (mask.view(n, 1, 1, c) if variant == "mask" else (f if variant == "gate" else 1.0))
    if variant == "relu_z":
        gt = gt * (z > 0).float()
    if variant == "gate":
        s = torch.sigmoid(alpha * a32 + beta)
        gt = gt * s * (1 - s)
    xhat = (a32 - mean) * invstd
    s1, s2 = gt.double().sum((0, 1, 2)), (gt * xhat).double().sum((0, 1, 2))
    assert torch.allclose(r1[0][:c], s1, rtol=1e-4, atol=1e-3) and torch.allclose(r1[0][c:], s2, rtol=1e-4, atol=1e-3)


@pytest.mark.parametrize("geom", [(2, 32, 40, 32, 32, 3), (1, 24, 24, 64, 64, 3), (2, 20, 28, 32, 96, 1), (1, 16, 16, 128, 256, 3),
                                  (2, 33, 21, 16, 8, 3)])
def test_conv_fp16_output_with_fused_statistics(dev, geom):
    """pmfb_conv_fwd with out_half: the stored fp16 tensor is the fp32 result rounded once, the fused statistics are the
    sums of the STORED values."""
    import pmf_b200
    from pmf_b200 import _lib as L
    from pmf_b200.engine import ConvParam, Engine, WeightCache
    n, h, w, ci, co, k = geom
    torch.manual_seed(sum(geom))
    x = torch.randn(n, h, w, ci, device=dev)
    wt = torch.nn.Parameter(torch.randn(co, ci, k, k, device=dev) * 0.1)
    bias = torch.nn.Parameter(torch.randn(co, device=dev))

    class P:
        mods = {}
    with pmf_b200.precision("tf32"):
        E = Engine(P(), dev, True, False, WeightCache())
        cp = ConvParam("c", wt, bias, 1, k // 2, 1)
        e = E.cache.get(cp, False, E.st)
        from pmf_b200._lib import ACT_LEAKY
        out = {}
        for half in (False, True):
            y = torch.empty(n, h, w, co, device=dev, dtype=torch.float16 if half else torch.float32)
            sums = torch.zeros(2 * co, device=dev, dtype=torch.float64)
            fused = E._conv_launch(x, ci, False, e["fwd"], co, cp.fwd_taps(), n, h, w, y, E._epi(beta1=e["bias"], act=ACT_LEAKY),
                                   bn_stats=sums, out_half=half)
            assert fused, "the halo kernel should fuse the statistics of this layer"
            out[half] = (y, sums)
    torch.cuda.synchronize()
    y32, s32 = out[False]
    y16, s16 = out[True]
    assert torch.equal(y16, y32.half())
    st = y16.double()
    assert torch.allclose(s16[:co], st.sum((0, 1, 2)), rtol=1e-6, atol=1e-4)
    assert torch.allclose(s16[co:], (st * st).sum((0, 1, 2)), rtol=1e-6, atol=1e-4)


@pytest.mark.parametrize("shape", [(2, 24, 40, 32), (3, 17, 23, 64), (1, 9, 11, 8), (2, 30, 40, 256)])
def test_fused_bn_finalisation_equals_the_separate_launch(dev, shape):
    """pmfb_pointwise16_bn (finalisation inside the BN-apply launch) == pmfb_bn_finalize followed by pmfb_pointwise16, bit
    for bit: outputs, alpha / beta / mean / invstd and the running statistics."""
    from pmf_b200 import _lib as L
    from pmf_b200._lib import BnFuse, Epilogue
    from pmf_b200.engine import _view
    n, h, w, c = shape
    st = torch.cuda.current_stream().cuda_stream
    g = torch.Generator(device="cpu").manual_seed(sum(shape))
    a16 = (torch.randn(n, h, w, c, generator=g) * 2 + 0.3).half().to(dev)
    ad = a16.double()
    sums = torch.cat([ad.sum((0, 1, 2)), (ad * ad).sum((0, 1, 2))]).contiguous()
    gamma = (torch.rand(c, generator=g) + 0.5).to(dev)
    beta = torch.randn(c, generator=g).to(dev)
    shortcut = torch.randn(n, h, w, c, generator=g).to(dev)
    res = []
    for fused in (False, True):
        rm, rv = torch.zeros(c, device=dev), torch.ones(c, device=dev)
        v = torch.empty(4 * c, device=dev)
        alpha, bt, mean, invstd = (v[i * c:(i + 1) * c] for i in range(4))
        y = torch.empty(n, h, w, c, device=dev)
        y16 = torch.empty(n, h, w, c, device=dev, dtype=torch.float16)
        e = Epilogue()
        e.r1, e.round_out = _view(shortcut), 1
        if fused:
            f = BnFuse()
            f.sums, f.count, f.gamma, f.beta = sums.data_ptr(), n * h * w, gamma.data_ptr(), beta.data_ptr()
            f.running_mean, f.running_var, f.momentum, f.eps = rm.data_ptr(), rv.data_ptr(), 0.1, 1e-5
            f.alpha_out, f.beta_out, f.mean_out, f.invstd_out = alpha.data_ptr(), bt.data_ptr(), mean.data_ptr(), invstd.data_ptr()
            L.call("pmfb_pointwise16_bn", C.byref(_view(a16)), y.data_ptr(), c * h * w, c * w, c, n, h, w, c, C.byref(e),
                   y16.data_ptr(), L.DT_F16, None, 1, C.byref(f), st)
        else:
            L.call("pmfb_bn_finalize", sums.data_ptr(), n * h * w, c, gamma.data_ptr(), beta.data_ptr(), rm.data_ptr(), rv.data_ptr(),
                   0.1, 1e-5, alpha.data_ptr(), bt.data_ptr(), mean.data_ptr(), invstd.data_ptr(), st)
            e.alpha1, e.beta1 = alpha.data_ptr(), bt.data_ptr()
            L.call("pmfb_pointwise16", C.byref(_view(a16)), y.data_ptr(), c * h * w, c * w, c, n, h, w, c, C.byref(e),
                   y16.data_ptr(), L.DT_F16, None, 1, st)
        torch.cuda.synchronize()
        res.append((y, y16, v, rm, rv))
    for t0, t1 in zip(*res):
        assert torch.equal(t0, t1)

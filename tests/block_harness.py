"""Shared block-level parity harness: runs one pmf_b200.net block through pmf_b200.engine (forward, then backward
from a seeded output gradient) and the same block through the oracle + torch.autograd, and returns the errors.
Used on the CPU box with the numpy C-ABI model (tests/cabi_mock.py) and on the B200 with the real library."""
import numpy as np
import torch

from oracle import pmf_oracle as po
from pmf_b200 import modules as M
from pmf_b200 import net as G
from pmf_b200.engine import Act, Engine, WeightCache


def _rel(a, b):
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-12))


def _l2(a, b, floor=1e-12):
    return float((a - b).double().norm() / max(float(b.double().norm()), floor))


def mask(rs, n, c, keep=0.7):
    return torch.from_numpy(((rs.rand(n, c, 1, 1) < keep) / keep).astype(np.float32))


def run_block(device, mod, build, oracle_fn, inputs, masks=None, train=True, multi=False, tf32=False, seed=3, in_kw=None):
    """Returns dict(fwd=[...], dinput=[...], dparam={name: err}, stats={name: err})."""
    shapes = {k: tuple(v.shape) for k, v in mod.state_dict().items()}
    sd = po.synth_state_dict(shapes, seed=seed)
    mod.load_state_dict(sd)
    mod.to(device)
    mod.train(train)
    masks = masks or {}
    rs = np.random.RandomState(seed)
    E = Engine(G.ModuleParams(mod), torch.device(device), train, train, WeightCache(),
               dropout={k: v.reshape(v.shape[0], -1) for k, v in masks.items()} if masks else False)
    in_kw = in_kw or [dict(c_pad=x.shape[1], needs_grad=True) for x in inputs]
    acts = [E.input_nchw(x.to(device), rnd=tf32, **kw) for x, kw in zip(inputs, in_kw)]
    ys = build(E, *acts)
    ys = ys if multi else (ys,)
    E.finish_forward()
    outs = [E.to_nchw(y.t, y.c).cpu() for y in ys]
    params = {"b." + k: (v.clone().requires_grad_(True) if v.dtype.is_floating_point and "running" not in k else v.clone())
              for k, v in sd.items()}
    xin = [x.clone().requires_grad_(True) for x in inputs]
    from pmf_b200 import engine as _eng
    ctx = po.Ctx(params, train=train, dropout={"b" + k: v for k, v in masks.items()}, tf32=tf32,
                 half_pre_bn=bool(E.h16 and _eng.PRE_BN_HALF))
    refs = oracle_fn(ctx, *xin)
    refs = refs if multi else (refs,)
    res = dict(fwd=[_rel(o, r.detach()) for o, r in zip(outs, refs)], dinput=[], dparam={}, stats={}, dinput_l2=[],
               dparam_l2={})
    if not train:
        return res
    total = 0
    for y, ref in zip(ys, refs):
        dout = torch.from_numpy(rs.normal(0, 1, tuple(ref.shape)).astype(np.float32))
        gy, _ = y.grad_target()
        E.input_nchw(dout.to(device), y.c, out=Act(gy, needs_grad=False), rnd=False)
        total = total + (ref * dout).sum()
    grads = E.run_backward()
    total.backward()
    for a, x in zip(acts, xin):
        if a.needs_grad:
            g = E.to_nchw(a.grad_read(), a.c).cpu()
            res["dinput"].append(_rel(g, x.grad))
            res["dinput_l2"].append(_l2(g, x.grad))
    for n, _ in mod.named_parameters():
        r = params["b." + n].grad
        wn = "b." + n.rsplit(".", 1)[0] + ".weight"
        scale = 1e-2 * float(params[wn].grad.abs().max())
        res["dparam"][n] = float((grads[n].cpu() - r).abs().max() / max(float(r.abs().max()), scale))
        res["dparam_l2"][n] = _l2(grads[n].cpu(), r, 1e-2 * float(params[wn].grad.double().norm()))
    live = mod.state_dict()
    for k, v in ctx.new_stats.items():
        res["stats"][k] = _rel(live[k[2:]].cpu(), v)
    return res


def block_cases(rs):
    """[(name, module, build, oracle_fn, inputs, masks, multi, in_kw)] covering every block type of the PMF graph."""
    t = lambda *s: torch.from_numpy(rs.normal(0, 1, s).astype(np.float32))  # noqa: E731
    cases = []
    cases.append(("res_context", M.ResContextBlock(16, 32), lambda E, a: G.res_context_block(E, a, ""),
                  lambda c, a: po.res_context_block(c, a, "b"), [t(2, 16, 8, 16)], None, False))
    cases.append(("res_block_pool_drop", M.ResBlock(16, 32, 0.2, pooling=True), lambda E, a: G.res_block(E, a, "", pooling=True),
                  lambda c, a: po.res_block(c, a, "b", pooling=True), [t(2, 16, 8, 16)], {".dropout": mask(rs, 2, 32)}, True))
    cases.append(("res_block_nopool_drop", M.ResBlock(16, 32, 0.2, pooling=False),
                  lambda E, a: G.res_block(E, a, "", pooling=False), lambda c, a: po.res_block(c, a, "b", pooling=False),
                  [t(2, 16, 8, 16)], {".dropout": mask(rs, 2, 32)}, False))
    cases.append(("up_block_drop", M.UpBlock(64, 16, 0.2), lambda E, a, s: G.up_block(E, a, s, ""),
                  lambda c, a, s: po.up_block(c, a, s, "b"), [t(2, 64, 4, 8), t(2, 32, 8, 16)],
                  {".dropout1": mask(rs, 2, 16), ".dropout2": mask(rs, 2, 48), ".dropout3": mask(rs, 2, 16)}, False))
    cases.append(("up_block_real_dims", M.UpBlock(128, 128, 0.2), lambda E, a, s: G.up_block(E, a, s, "", drop_out=False),
                  lambda c, a, s: po.up_block(c, a, s, "b", drop_out=False), [t(1, 128, 4, 8), t(1, 256, 8, 16)], None, False))
    cases.append(("aspp_tiny", M.ASPP(32, 32), lambda E, a: G.aspp(E, a, ""), lambda c, a: po.aspp(c, a, "b"),
                  [t(2, 32, 2, 4)], None, False))
    cases.append(("aspp_40x30", M.ASPP(32, 32), lambda E, a: G.aspp(E, a, ""), lambda c, a: po.aspp(c, a, "b"),
                  [t(1, 32, 30, 40)], None, False))

    cases = [c + (None,) for c in cases]
    from torchvision.models.resnet import BasicBlock, Bottleneck
    import torch.nn as nn

    def ds(cin, cout, stride):
        return nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))

    cases.append(("basic_block_s1", BasicBlock(32, 32), lambda E, a: G.basic_block(E, a, "", False),
                  lambda c, a: po.basic_block(c, a, "b", 1), [t(2, 32, 8, 16)], None, False, None))
    cases.append(("basic_block_s2_ds_drop", BasicBlock(32, 64, 2, ds(32, 64, 2)),
                  lambda E, a: G.basic_block(E, a, "", True, mask=E.mask_for(".drop", 2, 64)),
                  lambda c, a: c.drop(po.basic_block(c, a, "b", 2), "b.drop"), [t(2, 32, 16, 16)], {".drop": mask(rs, 2, 64)},
                  False, None))
    cases.append(("bottleneck_s2_ds", Bottleneck(64, 32, 2, ds(64, 128, 2)), lambda E, a: G.bottleneck(E, a, "", True),
                  lambda c, a: po.bottleneck(c, a, "b", 2), [t(2, 64, 8, 16)], None, False, None))

    class Stem(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv1 = nn.Conv2d(3, 64, kernel_size=7, stride=1, padding=3, bias=False)
            self.bn1 = nn.BatchNorm2d(64)

    def stem_build(E, a):
        return E.pool(E.conv_bn(a, ".conv1", ".bn1", post=G.ACT_RELU), "max")

    def stem_oracle(c, x):
        import torch.nn.functional as F
        return F.max_pool2d(F.relu(c.bn(c.conv(x, "b.conv1", padding=3), "b.bn1")), kernel_size=3, stride=2, padding=1)

    cases.append(("stem_maxpool", Stem(), stem_build, stem_oracle, [t(2, 3, 16, 24)], None, False,
                  [dict(c_pad=32, n_shift=7, needs_grad=False)]))

    class Dec(nn.Module):
        def __init__(self):
            super().__init__()
            self.camera_stream_decoder = M.RGBDecoder([16, 32, 64, 128], nclasses=20, base_channels=16)

    def dec_build(E, f0, f1, f2, f3):
        return G.rgb_decoder(E, [f0, f1, f2, f3], "camera_stream_decoder")

    def dec_oracle(c, f0, f1, f2, f3):
        # logits (pre-softmax) of RGBDecoder.forward
        import torch.nn.functional as F

        def stage(x, name, padding):
            x = c.bn(po.lrelu(c.conv(x, name + ".0", padding=padding)), name + ".2")
            return F.interpolate(x, scale_factor=2, mode="bilinear")

        p = "b.camera_stream_decoder"
        u4 = stage(f3, p + ".up_4a", 1)
        u3 = stage(torch.cat((u4, f2), 1), p + ".up_3a", 1)
        u2 = stage(torch.cat((u3, f1), 1), p + ".up_2a", 1)
        u1 = stage(torch.cat((u2, f0), 1), p + ".up_1a", 0)
        return c.conv(u1, p + ".conv", padding=1)

    cases.append(("rgb_decoder", Dec(), dec_build, dec_oracle,
                  [t(1, 16, 16, 32), t(1, 32, 8, 16), t(1, 64, 4, 8), t(1, 128, 2, 4)], None, False, None))
    return cases

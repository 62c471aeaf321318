#!/usr/bin/env python
"""Bring-up probe of the bf16 wgrad path (conv_wgrad_halo_kernel, kind::f16, MN-major 128B-swizzled operands): packed weight
gradients from bf16 shadows against the fp32 torch reference for a few stride-1 geometries."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pmf_b200 import _lib as L  # noqa: E402
from pmf_b200.engine import ConvParam, Engine, WeightCache, Act  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    L.require_device()
    torch.manual_seed(0)
    L.set_precision("f16")

    class P:
        mods = {}
    for (n, h, w, ci, co, k, dil) in ((2, 32, 40, 64, 64, 3, 1), (1, 24, 24, 128, 64, 3, 2), (2, 16, 24, 256, 256, 1, 1),
                                      (1, 40, 32, 64, 160, 3, 1), (1, 16, 16, 192, 64, 2, 2)):
        pad = dil * (k - 1) // 2 if k != 2 else 1
        x = torch.randn(n, h, w, ci, device=dev)
        dy = torch.randn(n, h, w, co, device=dev) * 0.1
        wt = torch.nn.Parameter(torch.randn(co, ci, k, k, device=dev) * 0.05)
        xr = x.permute(0, 3, 1, 2).detach().requires_grad_(False)
        out = torch.nn.functional.conv2d(xr, wt, padding=pad, dilation=dil)
        (out * dy.permute(0, 3, 1, 2)).sum().backward()
        ref = wt.grad
        res = {}
        for mode in ("tf32", "bf16"):
            E = Engine(P(), dev, True, True, WeightCache())
            E.use_side = False
            cp = ConvParam("c", wt, None, dil, pad, 1)
            xa = Act(x.clone(), needs_grad=False)
            E._dpre16 = dy.to(torch.bfloat16) if mode == "bf16" else None
            E.h16 = mode == "bf16"
            E._conv_bwd(xa, cp, dy.clone())
            torch.cuda.synchronize()
            g = E.param_grads["c.weight"]
            res[mode] = float((g - ref).abs().max() / ref.abs().max())
        print("wgrad n%d %dx%d ci%d co%d k%d d%d: max rel err tf32 %.2e  bf16 %.2e" % (n, h, w, ci, co, k, dil, res["tf32"], res["bf16"]))


if __name__ == "__main__":
    main()

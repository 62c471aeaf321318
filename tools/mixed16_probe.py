#!/usr/bin/env python
"""Hardware probe: does tcgen05.mma kind::f16 accept an fp16 A operand with a bf16 B operand (separate a_format / b_format
in the instruction descriptor)?  Runs one 3x3 64->64 convolution through pmfb_conv_fwd with x fp16 and w bf16 and compares
with the fp32 path.  (The wgrad of the "f16" mode wants x in fp16 and dy in bf16.)"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pmf_b200 import _lib as L  # noqa: E402
from pmf_b200.engine import ConvParam, Engine, WeightCache, _view  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    L.require_device()
    torch.manual_seed(0)
    n, h, w, ci, co = 2, 32, 40, 64, 64
    x = torch.randn(n, h, w, ci, device=dev)
    wt = torch.nn.Parameter(torch.randn(co, ci, 3, 3, device=dev) * 0.05)

    class P:
        mods = {}
    E = Engine(P(), dev, False, False, WeightCache())
    cp = ConvParam("c", wt, None, 1, 1, 1)
    e = E.cache.get(cp, False, E.st)
    ref = torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), wt, padding=1).permute(0, 2, 3, 1)
    out = {}
    for name, xt, wtp, dt in (("fp16xfp16", torch.float16, torch.float16, L.DT_F16), ("bf16xbf16", torch.bfloat16, torch.bfloat16, L.DT_BF16),
                              ("fp16xbf16", torch.float16, torch.bfloat16, 3)):
        y = torch.empty(n, h, w, co, device=dev)
        x16, w16 = x.to(xt), e["fwd"].to(wtp)
        E._conv_launch(x, ci, False, e["fwd"], co, cp.fwd_taps(), n, h, w, y, E._epi(), x16=x16, w16=w16, dt16=dt)
        torch.cuda.synchronize()
        out[name] = float((y - ref).abs().max() / ref.abs().max())
    print("max rel error vs fp32 conv:", out)


if __name__ == "__main__":
    main()

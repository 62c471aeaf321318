#!/usr/bin/env python
"""Micro-benchmark of the BatchNorm elementwise passes (BN apply = pmfb_pointwise16, backward reduce / apply) with the
pre-BN activation stored as fp32 vs fp16.  CUDA-event timing, inputs larger than L2.  Usage: python tools/elem_micro.py"""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from pmf_b200 import _lib as L  # noqa: E402
from pmf_b200._lib import Epilogue, View  # noqa: E402
from pmf_b200.engine import _view  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    dev = torch.device("cuda:0")
    L.require_device()
    st = torch.cuda.current_stream().cuda_stream
    only = sys.argv[1] if len(sys.argv) > 1 else ""
    for (n, h, w, c) in ((8, 480, 640, 32), (8, 240, 320, 64), (8, 120, 160, 128), (8, 60, 80, 256)):
        a32 = torch.randn(n, h, w, c, device=dev)
        a16 = a32.half()
        dy = torch.randn(n, h, w, c, device=dev)
        y = torch.empty(n, h, w, c, device=dev)
        y16 = torch.empty(n, h, w, c, device=dev, dtype=torch.float16)
        yb = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        dx = torch.empty(n, h, w, c, device=dev)
        dxb = torch.empty(n, h, w, c, device=dev, dtype=torch.bfloat16)
        vec = torch.rand(6 * c, device=dev) + 0.5
        alpha, beta, mean, invstd, gamma = (vec[i * c:(i + 1) * c] for i in range(5))
        red = torch.zeros(2 * c, device=dev, dtype=torch.float64)
        gw, gb = torch.empty(c, device=dev), torch.empty(c, device=dev)
        e = Epilogue()
        e.alpha1, e.beta1, e.round_out = alpha.data_ptr(), beta.data_ptr(), 1
        nv = View()
        px = n * h * w
        res = {}
        for half in (0, 1):
            a = a16 if half else a32
            av = _view(a)
            if only in ("", "pw"):
                res["pw_both%d" % half] = (timeit(lambda: L.call("pmfb_pointwise16", C.byref(av), y.data_ptr(), c * h * w, c * w, c, n, h, w, c,
                                                               C.byref(e), y16.data_ptr(), L.DT_F16, yb.data_ptr(), half, st)), (4 - 2 * half + 8) * c * px)
                res["pw_f32only%d" % half] = (timeit(lambda: L.call("pmfb_pointwise16", C.byref(av), y.data_ptr(), c * h * w, c * w, c, n, h, w, c,
                                                                  C.byref(e), None, L.DT_F16, None, half, st)), (4 - 2 * half + 4) * c * px)
            if only in ("", "red"):
                res["reduce%d" % half] = (timeit(lambda: L.call("pmfb_bn_bwd_reduce16", C.byref(_view(dy)), C.byref(nv), C.byref(nv), 0, C.byref(av),
                                                              mean.data_ptr(), invstd.data_ptr(), alpha.data_ptr(), beta.data_ptr(), n, h, w, c,
                                                              red.data_ptr(), half, st)), (8 - 2 * half) * c * px)
            if only in ("", "app"):
                for name, d32, d16, ob in (("apply_f32", dx, None, 4), ("apply_bf16", None, dxb, 2)):
                    res["%s%d" % (name, half)] = (timeit(lambda: L.call(
                        "pmfb_bn_bwd_apply16", C.byref(_view(dy)), C.byref(nv), C.byref(nv), 0, C.byref(av), mean.data_ptr(), invstd.data_ptr(),
                        alpha.data_ptr(), beta.data_ptr(), gamma.data_ptr(), red.data_ptr(), 1, n, h, w, c,
                        None if d32 is None else d32.data_ptr(), c * h * w, c * w, c, 1, gw.data_ptr(), gb.data_ptr(), None, None, 0, 0, 0, 0,
                        None if d16 is None else d16.data_ptr(), half, st)), (8 - 2 * half + ob) * c * px)
        print("shape", (n, h, w, c), " ".join("%s=%.3fms/%.0fGB/s" % (k, ms, by / ms / 1e6) for k, (ms, by) in res.items()), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Micro-benchmark of single convolution launches through the C-ABI (CUDA events, L2-cold: the tensors of one launch
exceed L2 at the default shapes).  Env knobs of the library (PMFB_*) are read once per process, so run one process per
setting:   PMFB_HALO_TMA_STORE=0 python tools/conv_micro.py
Shapes: (c_in, c_out, k, dilation, H, W) at batch 8; modes: train epilogue (bias+LeakyReLU [+fused BN stats]) and the eval
fusion (bias+LeakyReLU+BN affine+round)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn as nn

from pmf_b200 import net as G
from pmf_b200.engine import Engine, WeightCache

SHAPES = [(32, 32, 3, 1, 480, 640), (32, 32, 3, 2, 480, 640), (32, 32, 1, 1, 480, 640), (64, 64, 3, 1, 480, 640),
          (64, 64, 3, 1, 240, 320), (128, 128, 3, 1, 120, 160), (256, 256, 3, 1, 60, 80), (512, 512, 3, 1, 30, 40),
          (192, 64, 1, 1, 480, 640), (64, 192, 1, 1, 480, 640)]


class Blk(nn.Module):
    def __init__(self, ci, co, k, d):
        super().__init__()
        self.conv = nn.Conv2d(ci, co, k, padding=d * (k - 1) // 2, dilation=d)
        self.bn = nn.BatchNorm2d(co)


def main(B=8, reps=20):
    dev = torch.device("cuda:0")
    print("%-34s %10s %10s %10s" % ("shape", "train ms", "eval ms", "TFLOP/s(train)"))
    for (ci, co, k, d, H, W) in SHAPES:
        m = Blk(ci, co, k, d).to(dev)
        res = []
        for train in (True, False):
            m.train(train)
            E = Engine(G.ModuleParams(m), dev, train, False, WeightCache(), dropout=False)
            x = E.new(B, H, W, ci, needs_grad=False)
            x.t.normal_()
            for _ in range(3):
                E.conv_act_bn(x, "conv", "bn")
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            tot = 0.0
            for _ in range(reps):
                y = torch.empty((B, H, W, co), device=dev)
                a = torch.empty((B, H, W, co), device=dev)
                cp, bn = E.P.conv("conv"), E.P.bn("bn")
                e = E.cache.get(cp, False, E.st)
                if train:
                    sums = E.d64.take(2 * co)
                    e0.record()
                    E._conv_fwd(x, cp, a, E._epi(beta1=e["bias"], act=2), bn_stats=sums)
                    e1.record()
                else:
                    alpha, beta = E._bn_eval_affine(bn)
                    e0.record()
                    E._conv_fwd(x, cp, y, E._epi(beta1=e["bias"], act=2, alpha2=alpha, beta2=beta, rnd=1))
                    e1.record()
                torch.cuda.synchronize()
                tot += e0.elapsed_time(e1)
            res.append(tot / reps)
        fl = 2.0 * B * H * W * ci * co * k * k
        print("%-34s %10.3f %10.3f %10.1f" % ("%d->%d %dx%d d%d @%dx%d" % (ci, co, k, k, d, H, W), res[0], res[1],
                                               fl / (res[0] * 1e-3) / 1e12))


if __name__ == "__main__":
    main()

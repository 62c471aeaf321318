#!/usr/bin/env python
"""Kernel timeline of GRAPH-REPLAYED training steps (the mode bench.py times) from torch.profiler / CUPTI: per-kernel
totals, per-stream busy time, idle gaps on the main stream and the overlap of the side-stream weight-gradient kernels.
(The ncu launch lists under profiles/ are eager, serialised launches; this is the replayed step.)

    python tools/graph_timeline.py [--batch 8 --height 480 --width 640] > gpurun_out/timeline.txt
"""
import argparse
import collections
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from torch.profiler import ProfilerActivity, profile  # noqa: E402

import bench  # noqa: E402
import pmf_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--steps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    model = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34").to(dev).train()
    feat, label = bench.make_frames(a.batch, a.height, a.width, seed=1)
    x, y = feat.to(dev), label.to(dev)
    opt_a, opt_b = bench.make_optimizers(list(model.lidar_stream.parameters()),
                                         list(model.camera_stream_encoder.parameters()) +
                                         list(model.camera_stream_decoder.parameters()))
    from pmf_b200.loss import TrainerLoss
    crit = TrainerLoss(bench.NCLASSES, None, bench.LAMBDA, bench.GAMMA, bench.TAU, impl="auto").to(dev)

    def step():
        lid, cam = model(x[:, 0:5], x[:, 5:8])
        loss = crit(lid, cam, y)
        opt_a.zero_grad(set_to_none=True)
        opt_b.zero_grad(set_to_none=True)
        loss.backward()
        opt_a.step()
        opt_b.step()

    for _ in range(4):
        step()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        for _ in range(a.steps):
            step()
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA and e.time_range is not None]
    ks = sorted(((e.time_range.start, e.time_range.end, e.name, getattr(e, "device_index", 0)) for e in evs), key=lambda t: t[0])
    if not ks:
        print("no CUDA events captured")
        return
    t0, t1 = ks[0][0], max(k[1] for k in ks)
    print("steps %d  span %.3f ms  -> %.3f ms/step, %d kernels/step" % (a.steps, (t1 - t0) / 1e3, (t1 - t0) / 1e3 / a.steps, len(ks) // a.steps))
    tot = collections.Counter()
    cnt = collections.Counter()
    for s, e, n, _ in ks:
        n = n.split("(")[0].replace("void ", "").replace("pmfb::", "")[:70]
        tot[n] += e - s
        cnt[n] += 1
    print("\nkernel | launches/step | ms/step")
    for n, v in tot.most_common(28):
        print("%-72s %6.1f %8.3f" % (n, cnt[n] / a.steps, v / 1e3 / a.steps))
    # union of busy intervals (any stream) and the idle time inside the span
    busy, cur_s, cur_e = 0.0, None, None
    for s, e, _, _ in ks:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                busy += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    busy += cur_e - cur_s
    print("\nGPU busy (union over streams) %.3f ms/step, idle %.3f ms/step; sum of kernel durations %.3f ms/step"
          % (busy / 1e3 / a.steps, ((t1 - t0) - busy) / 1e3 / a.steps, sum(tot.values()) / 1e3 / a.steps))
    wg = [(s, e) for s, e, n, _ in ks if "wgrad" in n]
    other = [(s, e) for s, e, n, _ in ks if "wgrad" not in n]
    # time during which a wgrad kernel runs and no other kernel does
    import bisect
    other.sort()
    starts = [o[0] for o in other]
    alone = 0.0
    for s, e in wg:
        i = max(0, bisect.bisect_left(starts, s) - 2)
        cov = []
        while i < len(other) and other[i][0] < e:
            if other[i][1] > s:
                cov.append((max(s, other[i][0]), min(e, other[i][1])))
            i += 1
        c, last = 0.0, s
        for cs, ce in sorted(cov):
            if ce > last:
                c += ce - max(cs, last)
                last = ce
        alone += (e - s) - c
    print("weight-gradient kernels: %.3f ms/step in total, of which %.3f ms/step with no other kernel running" %
          (sum(e - s for s, e in wg) / 1e3 / a.steps, alone / 1e3 / a.steps))


if __name__ == "__main__":
    main()

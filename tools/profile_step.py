#!/usr/bin/env python
"""One eagerly launched (CUDA graphs off) PMF-ResNet34 training step between cudaProfilerStart/Stop, for
`ncu --profile-from-start off ...`.  Same workload as bench.py (batch 8, 480x640) unless overridden.

    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv \
        --log-file gpurun_out/launches.csv python tools/profile_step.py
    ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:conv_fwd_halo \
        -c 24 -o gpurun_out/halo python tools/profile_step.py
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["PMFB_CUDA_GRAPH"] = "0"

import torch  # noqa: E402

import bench  # noqa: E402
import pmf_b200  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--eval", action="store_true", help="profile an eval-mode forward instead of a training step")
    ap.add_argument("--loss", default="fused", choices=["fused", "torch"])
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    torch.manual_seed(1)
    model = pmf_b200.PMFNet(5, 3, 20, 32, False, "resnet34").to(dev)
    feat, label = bench.make_frames(a.batch, a.height, a.width, seed=1)
    x, y = feat.to(dev), label.to(dev)
    if a.eval:
        model.eval()

        def step():
            with torch.no_grad():
                model(x[:, 0:5], x[:, 5:8])
    else:
        model.train()
        opt_a, opt_b = bench.make_optimizers(list(model.lidar_stream.parameters()),
                                             list(model.camera_stream_encoder.parameters()) +
                                             list(model.camera_stream_decoder.parameters()))

        from pmf_b200.loss import TrainerLoss
        crit = TrainerLoss(bench.NCLASSES, None, bench.LAMBDA, bench.GAMMA, bench.TAU,
                           impl="auto" if a.loss == "fused" else "torch").to(dev)

        def step():
            lid, cam = model(x[:, 0:5], x[:, 5:8])
            loss = crit(lid, cam, y)
            opt_a.zero_grad(set_to_none=True)
            opt_b.zero_grad(set_to_none=True)
            loss.backward()
            opt_a.step()
            opt_b.step()
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    step()
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Write a synthetic dataset in the SemanticKITTI directory layout the reference's loader reads
(pc_processor/dataset/semantic_kitti/parser.py:39-75, 143-174):

    <root>/<seq>/velodyne/NNNNNN.bin    float32 (N,4)  x y z intensity
    <root>/<seq>/labels/NNNNNN.label    uint32  (N,)   semantic id in the low 16 bits, instance id in the high 16
    <root>/<seq>/image_2/NNNNNN.png     8-bit RGB
    <root>/<seq>/calib.txt              "P0:".."P3:" 3x4 and "Tr:" 3x4 rows, 12 numbers each

so that the unchanged `tasks/pmf/main.py` (dataset classes hard-wired at trainer.py:101-137) can run end to end on
synthetic frames of the BASELINE shape: one organised 64x2048 sweep (tests/synth.lidar_sweep) + one 480x640 image per frame.
Labels are written as RAW SemanticKITTI ids (the public label convention; semantic-kitti.yaml `learning_map_inv`), the
loader maps them back to the 20 training classes.

    python tools/make_synthetic_kitti.py /data/synth_kitti --sequences 0 8 --frames 16
"""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from tests import synth  # noqa: E402

# training class -> raw SemanticKITTI id (semantic-kitti.yaml: learning_map_inv)
RAW_ID = np.array([0, 10, 11, 15, 18, 20, 30, 31, 32, 40, 44, 48, 49, 50, 51, 70, 71, 72, 80, 81], dtype=np.uint32)


def synthetic_image(H, W, seed):
    """Low-pass filtered uniform noise, uint8 RGB (the loader feeds raw [0,1] RGB: perspective_view_loader.py:95)."""
    rs = np.random.RandomState(seed)
    rgb = rs.uniform(0, 1, (H + 2, W + 2, 3))
    rgb = sum(rgb[i:i + H, j:j + W] for i in range(3) for j in range(3)) / 9.0
    return np.clip(np.round(rgb * 255.0), 0, 255).astype(np.uint8)


def write_calib(path, H, W):
    P2, Tr = synth.camera_calibration(H, W)
    with open(path, "w") as f:
        for name in ("P0", "P1", "P2", "P3"):
            f.write("%s: %s\n" % (name, " ".join("%.12e" % v for v in P2.reshape(-1))))
        f.write("Tr: %s\n" % " ".join("%.12e" % v for v in Tr[:3].reshape(-1)))


def write_sequence(root, seq, n_frames, H=480, W=640, rows=64, cols=2048, seed=1):
    """Returns the list of (points, train_labels) written, for checks."""
    from PIL import Image
    sdir = os.path.join(root, "%02d" % int(seq))
    for sub in ("velodyne", "labels", "image_2"):
        os.makedirs(os.path.join(sdir, sub), exist_ok=True)
    write_calib(os.path.join(sdir, "calib.txt"), H, W)
    out = []
    for f in range(n_frames):
        fseed = (int(seed) * 1000 + int(seq)) * 1000 + f
        pts, lab = synth.lidar_sweep(rows, cols, seed=fseed)
        name = "%06d" % f
        pts.astype(np.float32).tofile(os.path.join(sdir, "velodyne", name + ".bin"))
        RAW_ID[lab].astype(np.uint32).tofile(os.path.join(sdir, "labels", name + ".label"))  # instance id 0
        Image.fromarray(synthetic_image(H, W, fseed)).save(os.path.join(sdir, "image_2", name + ".png"))
        out.append((pts, lab))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("root")
    ap.add_argument("--sequences", type=int, nargs="+", default=[0, 8])
    ap.add_argument("--frames", type=int, default=8)
    ap.add_argument("--height", type=int, default=480)
    ap.add_argument("--width", type=int, default=640)
    ap.add_argument("--seed", type=int, default=1)
    a = ap.parse_args()
    for s in a.sequences:
        write_sequence(a.root, s, a.frames, a.height, a.width, seed=a.seed)
        print("wrote sequence %02d: %d frames" % (s, a.frames))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""Stage an UNMODIFIED copy of the reference's Python sources under baseline/_ref/ (git-ignored; it travels to the GPU
box with the gpurun snapshot, /root/reference does not).  Used ONLY as the comparator / acceptance harness:

  * bench.py `gpu_eager_baseline` and `--impl reference`: the reference's own pc_processor.models.PMFNet and its loss
    classes (eager PyTorch/cuDNN on the GPU, oneDNN on the host cores);
  * tests/test_dropin_gpu.py: the byte-identical tasks/pmf/{main,option,trainer}.py driven over our pc_processor shim.

Nothing under baseline/_ref is product code and nothing in pmf_b200/ or pc_processor/ imports it.  The reference has no
setup.py / pyproject (pip install is not applicable), so "install" is a plain copy of the two source directories.

    python tools/stage_reference.py [/root/reference]
"""
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
DEST = os.path.join(ROOT, "baseline", "_ref")
PARTS = ("pc_processor", os.path.join("tasks", "pmf"), os.path.join("tasks", "pmf_eval_semantickitti"),
         os.path.join("tasks", "pmf_eval_nuscenes", "infer.py"))


def stage(src="/root/reference"):
    """Returns True when baseline/_ref holds a copy afterwards."""
    if not os.path.isdir(os.path.join(src, "pc_processor")):
        return os.path.isdir(os.path.join(DEST, "pc_processor"))
    os.makedirs(DEST, exist_ok=True)
    for part in PARTS:
        s, d = os.path.join(src, part), os.path.join(DEST, part)
        if os.path.isdir(s):
            if os.path.isdir(d):
                shutil.rmtree(d)
            shutil.copytree(s, d, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
        elif os.path.isfile(s):
            os.makedirs(os.path.dirname(d), exist_ok=True)
            shutil.copy2(s, d)
    with open(os.path.join(DEST, "STAGED_FROM"), "w") as f:
        f.write(src + "\n")
    return True


if __name__ == "__main__":
    ok = stage(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
    print("baseline/_ref staged" if ok else "reference tree not found; nothing staged")

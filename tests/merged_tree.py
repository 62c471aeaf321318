"""Builds the deployment layout of INTEGRATION.md §1 in a scratch directory (TEST INFRASTRUCTURE):

    <tmp>/PMF/pc_processor/      our shim's files (__init__, models/, postproc/, dataset/{__init__,perspective_view_loader}.py)
                                 + symlinks to every OTHER file / directory of the reference's pc_processor
    <tmp>/PMF/pmf_b200/          symlink to this repository's package (libpmf_b200.so inside)
    <tmp>/PMF/tasks/pmf/         symlinks to the reference's byte-identical main.py / option.py / trainer.py
    <tmp>/PMF/stubs/             tests/stubs (tensorboardX, nuscenes, pyquaternion, prettytable: not installed here)

The reference tree is /root/reference in the build container and the staged copy baseline/_ref on the GPU box
(tools/stage_reference.py); nothing is copied, everything is a symlink.
"""
import os

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def reference_root():
    for cand in (os.environ.get("PMF_REFERENCE_ROOT"), "/root/reference", os.path.join(ROOT, "baseline", "_ref")):
        if cand and os.path.isdir(os.path.join(cand, "pc_processor", "models")) and os.path.isdir(os.path.join(cand, "tasks", "pmf")):
            return cand
    return None


def _overlay(ours, theirs, dst):
    """dst = directory whose entries are symlinks to `ours` where we ship the entry, else to `theirs`; directories that
    exist on both sides are merged recursively."""
    os.makedirs(dst, exist_ok=True)
    names = set(n for n in os.listdir(theirs) if n != "__pycache__")
    if ours and os.path.isdir(ours):
        names |= set(n for n in os.listdir(ours) if n != "__pycache__")
    for n in sorted(names):
        o = os.path.join(ours, n) if ours else None
        t = os.path.join(theirs, n)
        if o and os.path.isdir(o) and os.path.isdir(t):
            _overlay(o, t, os.path.join(dst, n))
        elif o and os.path.exists(o):
            os.symlink(o, os.path.join(dst, n))
        else:
            os.symlink(t, os.path.join(dst, n))


def build(tmp, ref=None):
    """Returns the path of the merged PMF/ tree."""
    ref = ref or reference_root()
    if ref is None:
        raise RuntimeError("no reference tree (/root/reference or baseline/_ref)")
    top = os.path.join(str(tmp), "PMF")
    os.makedirs(top)
    _overlay(os.path.join(ROOT, "pc_processor"), os.path.join(ref, "pc_processor"), os.path.join(top, "pc_processor"))
    os.symlink(os.path.join(ROOT, "pmf_b200"), os.path.join(top, "pmf_b200"))
    os.symlink(os.path.join(ROOT, "tests", "stubs"), os.path.join(top, "stubs"))
    tdir = os.path.join(top, "tasks", "pmf")
    os.makedirs(tdir)
    for n in os.listdir(os.path.join(ref, "tasks", "pmf")):
        if n != "__pycache__":
            os.symlink(os.path.join(ref, "tasks", "pmf", n), os.path.join(tdir, n))
    return top

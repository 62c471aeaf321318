"""Import the UNMODIFIED reference modules from /root/reference (build container only).

``import pc_processor`` of the reference fails here (tensorboardX / nuscenes-devkit / pyquaternion are not
installed), so a stub parent package with the reference's ``__path__`` is registered under
the alias ``ref_pc_processor`` and only the sub-packages the hot path needs are imported.  Nothing in the
``-m gpu`` tests, ``smoke()`` or ``bench.py`` may call this: /root/reference does not exist on the GPU box.
"""
import importlib
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("PMF_REFERENCE_ROOT", "/root/reference")
ALIAS = "ref_pc_processor"


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "pc_processor", "models"))


def load_reference():
    """Returns a namespace with .models / .postproc / .loss / .metrics of the real reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    if ALIAS not in sys.modules:
        pkg = types.ModuleType(ALIAS)
        pkg.__path__ = [os.path.join(REFERENCE_ROOT, "pc_processor")]
        pkg.__package__ = ALIAS
        sys.modules[ALIAS] = pkg
    pkg = sys.modules[ALIAS]
    for sub in ("models", "postproc", "loss", "metrics"):
        setattr(pkg, sub, importlib.import_module(ALIAS + "." + sub))
    return pkg


def load_reference_file(relpath: str, name: str):
    """Import one reference source file by path (e.g. the dataset loaders, which need stubs otherwise)."""
    import importlib.util

    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, relpath))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

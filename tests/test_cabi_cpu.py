"""The C-ABI library builds, loads and exports exactly what include/pmfb.h declares (no compute calls: no GPU here),
and the ctypes binding agrees with the header's prototypes."""
import os
import re
import subprocess

import pytest

from pmf_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _header_prototypes():
    h = open(os.path.join(ROOT, "include", "pmfb.h")).read()
    h = re.sub(r"/\*.*?\*/", "", h, flags=re.S)
    return re.findall(r"(?:int|size_t|const char\*)\s+(pmfb_\w+)\s*\(([^;]*?)\)\s*;", h, flags=re.S)


def test_library_exports_every_declared_symbol():
    if not os.path.exists(_lib.LIB_PATH):
        import __graft_entry__
        __graft_entry__.build()
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    declared = {name for name, _ in _header_prototypes()}
    assert declared, "no prototypes parsed from include/pmfb.h"
    assert declared <= exported, sorted(declared - exported)
    assert {s for s in exported if s.startswith("pmfb_")} == declared


def test_binding_matches_header_prototypes():
    protos = dict(_header_prototypes())
    assert set(protos) == set(_lib.EXPORTS)
    for name, args in protos.items():
        args = args.strip()
        n = 0 if args in ("void", "") else len(args.split(","))
        assert n == len(_lib._SIGNATURES[name][0]), name


def test_library_loads_and_reports_abi_and_no_device():
    l = _lib.lib()
    assert l.pmfb_abi_version() == _lib.ABI_VERSION
    import torch
    if not torch.cuda.is_available():
        # no CPU fallback: without a device the library says so, and the binding raises
        assert l.pmfb_init() != 0
        assert "device" in _lib.last_error().lower() or "driver" in _lib.last_error().lower()
        with pytest.raises(_lib.PmfbError):
            _lib.require_device()


def test_struct_layout_matches_c():
    """sizeof() of the descriptor structs as the C compiler lays them out."""
    import ctypes as C
    src = r'''
    #include <stdio.h>
    #include "pmfb.h"
    #include <stddef.h>
    int main(){ printf("%zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu %zu\n", sizeof(pmfb_view), sizeof(pmfb_epilogue), sizeof(pmfb_tma_src),
                       sizeof(pmfb_conv_desc), sizeof(pmfb_wgrad_desc), sizeof(pmfb_weight_job),
                       offsetof(pmfb_weight_job, c_out), offsetof(pmfb_weight_job, start), offsetof(pmfb_conv_desc, bn_stats),
                       offsetof(pmfb_conv_desc, out_half), sizeof(pmfb_bn_fuse), offsetof(pmfb_bn_fuse, momentum),
                       offsetof(pmfb_bn_fuse, alpha_out));
                return 0; }
    '''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        open(os.path.join(d, "s.c"), "w").write(src)
        subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", os.path.join(d, "s"), os.path.join(d, "s.c")], check=True)
        sizes = [int(x) for x in subprocess.run([os.path.join(d, "s")], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_lib.View), C.sizeof(_lib.Epilogue), C.sizeof(_lib.TmaSrc), C.sizeof(_lib.ConvDesc),
                     C.sizeof(_lib.WgradDesc), C.sizeof(_lib.WeightJob), _lib.WeightJob.c_out.offset, _lib.WeightJob.start.offset,
                     _lib.ConvDesc.bn_stats.offset, _lib.ConvDesc.out_half.offset, C.sizeof(_lib.BnFuse), _lib.BnFuse.momentum.offset,
                     _lib.BnFuse.alpha_out.offset]


def test_bench_reference_arm_json_contract():
    """`bench.py --impl reference` (the reference's CPU path = the oracle port) prints ONE JSON line with the contract's
    keys; run on a tiny frame so that the CPU suite stays fast."""
    import json
    import sys
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1",
                          "--height", "32", "--width", "64"], capture_output=True, text=True, check=True, cwd=ROOT).stdout
    lines = [ln for ln in out.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    for k in ("metric", "n_gpus", "steps", "warmup", "ms_per_step", "scaling", "vs_baseline", "dtype", "data", "config"):
        assert k in d

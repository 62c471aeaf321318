#!/usr/bin/env python
"""Per-kernel SASS evidence of libpmf_b200.so (B200_PROFILING.md "What proves a Blackwell-native kernel"): counts of the
tcgen05 / TMEM / TMA mnemonics (UTC*MMA, LDTM, UTMALDG, UTMASTG, UTCBAR ...) per kernel from `cuobjdump -sass`, plus a few
instruction excerpts around the first UTC*MMA of the dominant kernels.

    python tools/sass_summary.py > profiles/r2_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pmf_b200", "libpmf_b200.so")
PAT = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "UTCOMMA", "UTCMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCBAR", "UTCATOMSWS", "UBLKCP", "SYNCS",
       "HMMA", "HGMMA", "REDG", "ATOMG", "ATOMS", "MATCH", "VOTE", "SHFL"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels = collections.OrderedDict()
    cur, arch = None, None
    for ln in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", ln)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("pmfb::", "")
            kernels[cur] = {"n": 0, "ops": collections.Counter(), "lines": []}
            continue
        m = re.match(r"\s*arch = (\S+)", ln)
        if m:
            arch = m.group(1)
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(.*?);", ln)
        if m:
            ins = m.group(1)
            k = kernels[cur]
            k["n"] += 1
            k["lines"].append(ins)
            for p in PAT:
                if re.search(r"\b%s" % p, ins):
                    k["ops"][p] += 1
    print("SASS summary of pmf_b200/libpmf_b200.so (cuobjdump -sass; arch %s; %d kernels)" % (arch, len(kernels)))
    tot = collections.Counter()
    for k in kernels.values():
        tot.update(k["ops"])
    print("totals: " + ", ".join("%s %d" % (p, tot[p]) for p in PAT if tot[p]))
    print()
    print("kernel | instructions | " + " | ".join(PAT[:12]))
    for name, k in kernels.items():
        if sum(k["ops"][p] for p in PAT[:12]) == 0:
            continue
        print("%s | %d | %s" % (name[:88], k["n"], " | ".join(str(k["ops"][p]) for p in PAT[:12])))
    for want in ("conv_fwd_halo_kernel<25>", "conv_wgrad_halo_kernel", "knn_vote_points_kernel"):
        for name, k in kernels.items():
            if name.startswith(want):
                idx = next((i for i, l in enumerate(k["lines"]) if "UTCHMMA" in l), None)
                if idx is None:
                    idx = next((i for i, l in enumerate(k["lines"]) if "LDG" in l), None)
                if idx is None:
                    continue
                print("\n--- %s: instructions around the first tensor / memory instruction" % name)
                for l in k["lines"][max(0, idx - 6):idx + 10]:
                    print("    " + l)
                break


if __name__ == "__main__":
    sys.exit(main())

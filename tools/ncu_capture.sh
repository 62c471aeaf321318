#!/bin/bash
# ncu --set full captures of selected launches of one eager training step (tools/profile_step.py), exported on the
# GPU box to CSV (raw metrics + per-line source counters) so that only small text files travel back in gpurun_out/.
#   tools/ncu_capture.sh <name> <kernel-regex> <launch-skip> <launch-count> [profile_step args...]
set -u
name=$1; regex=$2; skip=$3; count=$4; shift 4
out=gpurun_out
mkdir -p $out /tmp/ncu
ncu --profile-from-start off --set full --clock-control none --import-source on -k "regex:$regex" \
    --launch-skip "$skip" --launch-count "$count" -f -o /tmp/ncu/$name python tools/profile_step.py "$@" > $out/${name}_ncu.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>> $out/${name}_ncu.log
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source sass 2>> $out/${name}_ncu.log | gzip -9 > $out/${name}_src_sass.csv.gz
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source cuda 2>> $out/${name}_ncu.log | gzip -9 > $out/${name}_src_cuda.csv.gz
ls -la /tmp/ncu/$name.ncu-rep $out/${name}_* >> $out/${name}_ncu.log

#!/bin/bash
# ncu --set full capture of ONE launch of a micro-benchmark (tools/conv_micro.py by default), exported to CSV on the box.
#   tools/ncu_micro.sh <name> <kernel-regex> <launch-skip> [command...]
set -u
name=$1; regex=$2; skip=$3; shift 3
out=gpurun_out
mkdir -p $out /tmp/ncu
if [ $# -eq 0 ]; then set -- python tools/conv_micro.py; fi
ncu --set full --clock-control none --import-source on -k "regex:$regex" --launch-skip "$skip" --launch-count 1 -f -o /tmp/ncu/$name "$@" > $out/${name}_ncu.log 2>&1
ncu -i /tmp/ncu/$name.ncu-rep --page raw --csv > $out/${name}_raw.csv 2>> $out/${name}_ncu.log
ncu -i /tmp/ncu/$name.ncu-rep --page source --csv --print-source sass 2>> $out/${name}_ncu.log | gzip -9 > $out/${name}_src_sass.csv.gz

#!/bin/bash
# ncu --set full capture of selected kernels -> gpurun_out/<name>.ncu-rep + a CSV of the raw page.
# usage: [SKIP=n] tools/ncu_metrics.sh <name> <kernel-regex> <count> <cmd...>
name=$1; regex=$2; count=$3; shift 3
ncu --set full --clock-control none --import-source on --kernel-name-base mangled -k "regex:$regex" -s ${SKIP:-0} -c $count -o gpurun_out/$name -f "$@" > gpurun_out/$name.log 2>&1
ncu -i gpurun_out/$name.ncu-rep --page raw --csv > gpurun_out/$name.raw.csv 2>/dev/null
tail -2 gpurun_out/$name.log

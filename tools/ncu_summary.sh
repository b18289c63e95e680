#!/bin/bash
# Runs on the GPU box: one ncu --set full capture, reduced to text/CSV summaries so that
# gpurun_out stays small.  usage: tools/ncu_summary.sh <tag> <kernel-regex> <count> -- <command...>
set -u
tag=$1; regex=$2; count=$3; shift 4
rep=/tmp/prof_$tag
"$@" > gpurun_out/plain_$tag.log 2>&1 || { echo "plain run failed"; exit 1; }
ncu --set full --clock-control none -k "regex:$regex" -c "$count" -o $rep "$@" > gpurun_out/ncu_$tag.log 2>&1
ncu -i $rep.ncu-rep --page details 2>/dev/null | grep -vE '^\s*$' > gpurun_out/ncu_details_$tag.txt
ncu -i $rep.ncu-rep --page raw --csv 2>/dev/null > gpurun_out/ncu_raw_$tag.csv
rm -f $rep.ncu-rep
wc -c gpurun_out/ncu_details_$tag.txt gpurun_out/ncu_raw_$tag.csv

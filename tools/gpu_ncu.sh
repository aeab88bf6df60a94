#!/usr/bin/env bash
# tools/gpu_ncu.sh <tag> <kernel-regex> <workload...> -- one ncu --set full capture (with source) per workload
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG="$1"; KRE="$2"; shift 2
for WL in "$@"; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 6 -c 1 -f -o $OUT/prof_${TAG}_${WL} python bench.py --workload $WL --steps 1 --warmup 3 --skip-e2e > $OUT/ncu_${TAG}_${WL}.log 2>&1
done
ls -la $OUT | tail -4

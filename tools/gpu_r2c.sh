#!/usr/bin/env bash
# tools/gpu_r2c.sh <tag>[:wl1,wl2,...] <pytest args...> -- selected GPU tests, then kernel-side bench lines of the workloads
set -u
OUT=gpurun_out
mkdir -p $OUT
SPEC="$1"; shift
TAG="${SPEC%%:*}"; WLS="cfg3a,cfg3b"; [[ "$SPEC" == *:* ]] && WLS="${SPEC#*:}"
if [ $# -gt 0 ]; then timeout 1500 python -m pytest "$@" -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.txt; fi
for w in ${WLS//,/ }; do
  timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | sed "s/^/$w /" | tee -a $OUT/quick_$TAG.log
done

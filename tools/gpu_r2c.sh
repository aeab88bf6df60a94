#!/usr/bin/env bash
# tools/gpu_r2c.sh <tag> <pytest args...> -- selected GPU tests, then kernel-side bench lines of the workloads in $WLS
set -u
OUT=gpurun_out
mkdir -p $OUT
TAG="$1"; shift
timeout 1500 python -m pytest "$@" -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.txt
for w in ${WLS:-cfg3a cfg3b}; do
  timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1 | tee -a $OUT/quick_$TAG.log
done

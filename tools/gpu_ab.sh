#!/usr/bin/env bash
# tools/gpu_ab.sh <tag> <variant> <wl,wl,...> [pytest args] -- kernel-side bench lines of the in-tree library and of build/variants/libvgl_<variant>.so, interleaved
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG="$1"; VAR="$2"; WLS="$3"; shift 3
if [ $# -gt 0 ]; then timeout 1500 python -m pytest "$@" -x -q 2>&1 | tail -15 | tee $OUT/pytest_$TAG.txt; fi
for rep in 1 2; do
for w in ${WLS//,/ }; do
  timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-configs 2>&1 | tail -1 | sed "s/^/new $w /" | tee -a $OUT/ab_$TAG.log
  VGL_LIB=$PWD/build/variants/libvgl_$VAR.so timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-configs 2>&1 | tail -1 | sed "s/^/$VAR $w /" | tee -a $OUT/ab_$TAG.log
done
done

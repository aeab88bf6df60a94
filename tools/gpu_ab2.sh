#!/usr/bin/env bash
# tools/gpu_ab2.sh <tag> <wl,wl,...> <variant> [variant ...] -- kernel-side bench lines of the in-tree library and of each build/variants/libvgl_<variant>.so, interleaved
set -u
OUT=gpurun_out; mkdir -p $OUT
TAG="$1"; WLS="$2"; shift 2
for rep in 1 2; do
for w in ${WLS//,/ }; do
  timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-configs 2>&1 | tail -1 | sed "s/^/base $w /" | tee -a $OUT/ab_$TAG.log
  for v in "$@"; do
    VGL_LIB=$PWD/build/variants/libvgl_$v.so timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline --no-configs 2>&1 | tail -1 | sed "s/^/$v $w /" | tee -a $OUT/ab_$TAG.log
  done
done
done

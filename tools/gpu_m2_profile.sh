#!/usr/bin/env bash
# tools/gpu_m2_profile.sh -- run under gpurun: ncu --set full of k_tile_m2 on the cfg3 workloads + the new pileup test
set -u
TAG="${1:-r02d}"
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests/test_gpu_pileup.py -x -q 2>&1 | tail -3
for WL in cfg3a cfg3b; do
ncu --set full --clock-control none --import-source on -k regex:k_tile_m2 -s 5 -c 1 -f -o $OUT/prof_m2_${WL}_$TAG python bench.py --workload $WL --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_m2_${WL}_$TAG.log 2>&1
done
ls -la $OUT | tail -4

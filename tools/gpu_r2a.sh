#!/usr/bin/env bash
# tools/gpu_r2a.sh -- round 2 baseline: ncu --set full (with source) of the AUX tile kernel on cfg4 and of k_tile_m1f on cfg2
set -u
OUT=gpurun_out
mkdir -p $OUT
for WL in cfg4 cfg2; do
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_tile_m1f -s 5 -c 1 -f -o $OUT/prof_r2a_${WL} python bench.py --workload $WL --steps 2 --warmup 3 --skip-e2e > $OUT/ncu_r2a_${WL}.log 2>&1
done
ls -la $OUT | tail -4

#!/usr/bin/env bash
# tools/gpu_m2.sh -- run under gpurun: model-2 tile kernel tests + kernel-side bench of the cfg3 workloads
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_tile_m2.py -x -q 2>&1 | tail -15
for w in cfg3a cfg3b; do
  echo "== $w: $(timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1)" | tee -a gpurun_out/m2.log
done

#!/usr/bin/env bash
# tools/gpu_quick.sh -- run under gpurun: fused/selftest GPU tests + kernel-side bench of cfg2 and the cfg5 shape
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_selftest.py -x -q 2>&1 | tail -2
for w in cfg2 cfg5; do
  echo "== $w: $(python bench.py --workload $w --steps 8 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1)" | tee -a gpurun_out/quick.log
done

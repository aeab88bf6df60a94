#!/usr/bin/env bash
# tools/gpu_aux.sh -- run under gpurun: AUX tile kernel tests + cfg4 bench + regression of the other tile tests
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tile_aux.py -x -q 2>&1 | tail -25
timeout 900 python -m pytest tests/test_gpu_fused.py tests/test_gpu_narrow.py tests/test_gpu_native.py -x -q 2>&1 | tail -5
for w in cfg4 cfg2; do
  echo "== $w: $(timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1)" | tee -a gpurun_out/aux.log
done

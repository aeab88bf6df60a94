#!/usr/bin/env bash
# tools/gpu_variants.sh -- run under gpurun: kernel-side bench of every build/variants/libvgl_*.so (A/B of tunables)
mkdir -p gpurun_out
for so in build/variants/libvgl_*.so; do
  name=$(basename $so .so)
  echo "== $name: $(VGL_LIB=$PWD/$so python bench.py --steps 8 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1)" | tee -a gpurun_out/variants.log
done

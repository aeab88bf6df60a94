mkdir -p gpurun_out
for w in cfg4; do
  echo "== $w: $(timeout 600 python bench.py --workload $w --steps 4 --warmup 3 --skip-e2e --no-cpu-baseline 2>&1 | tail -1)" | tee -a gpurun_out/quick.log
done

#!/usr/bin/env bash
# tools/gpu_scale8.sh -- run under `gpurun --gpus 8`: the driver's N=4 and N=8 launches of the bench (cfg2) and the
# scaling config (cfg5 shape: 10 000 samples, depth 30) at N=8
mkdir -p gpurun_out
port=29530
for n in 4 8; do
  port=$((port+1))
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  tail -c 600 gpurun_out/bench_n$n.json | head -c 300; echo; tail -2 gpurun_out/bench_n$n.err
done
port=$((port+1))
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port bench.py --gpus 8 --steps 8 --warmup 3 --no-cpu-baseline --workload cfg5 > gpurun_out/bench_cfg5_n8.json 2> gpurun_out/bench_cfg5_n8.err
head -c 400 gpurun_out/bench_cfg5_n8.json; echo; tail -2 gpurun_out/bench_cfg5_n8.err

#!/usr/bin/env bash
# tools/gpu_scale.sh <N> <tag> [bench args] -- the driver's multi-GPU launch (default: its --steps 20 --warmup 5)
N="$1"; TAG="$2"; shift 2; OUT=gpurun_out; mkdir -p $OUT
ARGS="${*:---steps 20 --warmup 5}"
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N $ARGS > $OUT/bench_${TAG}_n$N.json 2> $OUT/bench_${TAG}_n$N.err; echo "rc=$?"; tail -c 400 $OUT/bench_${TAG}_n$N.err
python - <<P
import json
d=json.loads(open("$OUT/bench_${TAG}_n$N.json").read().strip().splitlines()[-1])
print("N=$N value %.4g per-gpu %.4g frac %.3f ms/step %.2f e2e %.4g"%(d["value"],d["value"]/$N,d["roofline"]["frac"],d["ms_per_step"],d["e2e"]["value"]))
P

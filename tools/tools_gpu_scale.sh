#!/bin/bash
# weak-scaling check on one box: N = 8 (and 4) ranks of bench.py over NCCL
mkdir -p gpurun_out
nvidia-smi -L | wc -l
for n in 8; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 200 --warmup 50 --cpu-seconds 1 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
  python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n$n.json') if l.startswith('{')][-1]); print($n, d['value'], d['ms_per_step'], d['e2e']['value'], d['outputs_finite'])"
done

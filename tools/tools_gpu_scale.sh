#!/bin/bash
# weak-scaling check on one box: N ranks of bench.py over NCCL (cfg2 default line, cfg4 strong scaling, cfg5 epoch sweep)
mkdir -p gpurun_out
N=${1:-8}
nvidia-smi -L | wc -l
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 200 --warmup 50 --cpu-seconds 1 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
python -c "
import json; d=json.loads([l for l in open('gpurun_out/bench_n$N.json') if l.startswith('{')][-1]); print($N, d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('frac'), d['outputs_finite'])"
for w in cfg4 cfg5; do
  timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29700+N)) bench.py --gpus $N --workload $w --steps 100 --warmup 20 2>> gpurun_out/bench_n$N.err | tail -1 > gpurun_out/bench_${w}_n$N.json
  python -c "
import json; d=json.loads(open('gpurun_out/bench_${w}_n$N.json').read()); print('$w', $N, d['value'], d['ms_per_step'], d.get('sweep_seconds'))"
done

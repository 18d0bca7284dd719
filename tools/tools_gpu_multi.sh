#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 20 --cpu-seconds 2 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; cat gpurun_out/bench_n2.json; tail -3 gpurun_out/bench_n2.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --impl reference --steps 5 --warmup 1 2>> gpurun_out/bench_n2.err | tail -1 | cut -c1-200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --workload cfg4 --steps 50 --warmup 10 2>> gpurun_out/bench_n2.err | tail -1 > gpurun_out/bench_cfg4_n2.json; cat gpurun_out/bench_cfg4_n2.json
timeout 300 python bench.py --workload cfg4 --steps 50 --warmup 10 > gpurun_out/bench_cfg4_n1.json 2>> gpurun_out/bench_n2.err; cat gpurun_out/bench_cfg4_n1.json
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 10 > gpurun_out/bench_cfg3_n1.json 2>> gpurun_out/bench_n2.err; cat gpurun_out/bench_cfg3_n1.json
timeout 600 python -m pytest tests -q -m gpu -k "cfg4 or cfg5" 2>&1 | tail -3

#!/bin/bash
# quick GPU pass: parity tests + bench (+ optional ncu capture of kernel regex $1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --cpu-seconds 2 > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'], d['e2e']['value'], d['clocks'])"; tail -5 gpurun_out/bench.err
if [ -n "$1" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:$1 -s 3 -c 1 -o gpurun_out/prof_$1 python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log
fi

#!/bin/bash
# first GPU pass: parity tests, smoke, bench, ncu launch list + one full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:foa_features -s 3 -c 1 -o gpurun_out/prof_foa python bench.py --steps 3 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out

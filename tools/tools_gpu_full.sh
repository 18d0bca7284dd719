#!/bin/bash
# Round-end validation on one B200: GPU test-suite, smoke, both bench arms, the extra workloads, ncu launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
for w in cfg3 cfg4 cfg5 wav2img augment; do
  timeout 300 python bench.py --workload $w --steps 200 --warmup 20 > gpurun_out/bench_$w.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_$w.json
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --cpu-seconds 0 > gpurun_out/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:scalar_wav2img -s 5 -c 1 -o gpurun_out/prof_epi python bench.py --workload wav2img --steps 10 --warmup 3 > gpurun_out/ncu_epi.log 2>&1
# ncu --set full of the two feature kernels (one launch each, after warm-up) -> tools/ncu_summary.py / tools_sass_hot.py read them here
timeout 400 ncu --set full --clock-control none --import-source on -k regex:foa_iv2 -s 5 -c 1 -f -o gpurun_out/prof_foa_r02 python tools/prof_foa.py > gpurun_out/ncu_foa.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:mic_features -s 5 -c 1 -f -o gpurun_out/prof_mic_r02 python tools/prof_foa.py mic > gpurun_out/ncu_mic.log 2>&1

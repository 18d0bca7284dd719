#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -12 gpurun_out/pytest_gpu.log
for cfg in "3 6" "3 8" "3 9" "3 10" "2 8"; do
  set -- $cfg
  echo "== iv kernel $1 pairs/warps $2"; SELD_IV_KERNEL=$1 SELD_IV3_PAIRS=$2 SELD_IV2_WARPS=$2 timeout 300 python bench.py --cpu-seconds 0 --steps 30 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])"
done

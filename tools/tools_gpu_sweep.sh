#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu.log 2>&1; tail -4 gpurun_out/pytest_gpu.log
for cfg in $SWEEP; do
  k=${cfg%%:*}; w=${cfg##*:}
  echo "== iv kernel $k pairs/warps $w"; SELD_IV_KERNEL=$k SELD_IV3_PAIRS=$w SELD_IV2_WARPS=$w timeout 300 python bench.py --cpu-seconds 0 --steps 30 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])"
done

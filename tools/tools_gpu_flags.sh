#!/bin/bash
for fl in $FLAGS; do
  echo "== SELD_FLAGS=$fl"; SELD_FLAGS=$fl timeout 300 python bench.py --cpu-seconds 0 --steps 40 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['roofline']['frac'])"
done

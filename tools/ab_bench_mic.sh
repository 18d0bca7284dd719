#!/bin/bash
for i in 1 2; do for v in A B; do
  echo -n "$v: "; SELD_LIB=$PWD/build/ab/lib$v.so timeout 60 python bench.py --workload cfg3 --steps 100 --warmup 20 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us  %.4f' % (1e3*d['ms_per_step'], d['roofline']['frac']))"
done; done

#!/bin/bash
# bench several builds under build/ab/ on the same box: ./tools/ab_multi.sh A P1 P2 ...
for i in 1 2; do for v in "$@"; do
  echo -n "$v: "; SELD_LIB=$PWD/build/ab/lib$v.so timeout 60 python bench.py --steps 200 --warmup 50 --cpu-seconds 0 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us  %.4f' % (1e3*d['ms_per_step'], d['roofline']['frac']))"
done; done

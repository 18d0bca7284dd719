#!/bin/bash
# A/B two builds of libseldfeat on the same box: build/ab/libA.so vs libB.so (alternating, 3 rounds)
for i in 1 2 3; do for v in A B; do
  echo -n "$v: "; SELD_LIB=$PWD/build/ab/lib$v.so timeout 60 python bench.py --steps 200 --warmup 50 --cpu-seconds 0 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('%.1f us  %.4f' % (1e3*d['ms_per_step'], d['roofline']['frac']))"
done; done

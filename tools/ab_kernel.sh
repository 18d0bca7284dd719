#!/bin/bash
# A/B the FOA kernels of ONE build on the same box: SELD_IV_KERNEL=5 (tensor-core mel) vs 2 (fp32 mel walk), alternating
for i in 1 2 3; do for v in ${KERNELS:-5 2}; do
  echo -n "iv$v: "; SELD_IV_KERNEL=$v timeout 120 python bench.py --steps 200 --warmup 50 --cpu-seconds 0 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('%.1f us  %.4f' % (1e3*d['ms_per_step'], d['roofline']['frac']))"
done; done

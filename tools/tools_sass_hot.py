#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass` dump: dynamic instruction mix,
shared-memory wavefronts and stall samples by opcode and by code region (address ranges)."""
import csv, sys, collections
rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[1]
ix = {h: i for i, h in enumerate(hdr)}
data = rows[2:]
tot_inst = 0; byop = collections.Counter(); wf = collections.Counter(); wfi = collections.Counter(); samp = collections.Counter()
NF = float(sys.argv[2]) if len(sys.argv) > 2 else 64064.0
seq = []
for r in data:
    if len(r) < len(hdr): continue
    src = r[ix['Source']].strip()
    toks = src.split()
    op = toks[0] if not toks[0].startswith('@') else toks[1]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS')) and '.' in op else '')
    n = int(r[ix['Instructions Executed']]); tot_inst += n; byop[op] += n
    w = int(r[ix['L1 Wavefronts Shared']] or 0); wi = int(r[ix['L1 Wavefronts Shared Ideal']] or 0)
    wf[op] += w; wfi[op] += wi
    samp[op] += int(r[ix['# Samples']] or 0)
    seq.append((r[ix['Address']], src, n, w, wi, int(r[ix['# Samples']] or 0)))
print('total warp-instr/frame %.0f' % (tot_inst / NF))
for op, n in byop.most_common(28):
    print('%-10s %8.1f /frame  smem wf %7.1f (ideal %7.1f)  samples %d' % (op, n / NF, wf[op] / NF, wfi[op] / NF, samp[op]))
# regions: split the instruction stream into 40 equal slices by position and print cumulative
if len(sys.argv) > 3:
    k = int(sys.argv[3]); step = (len(seq) + k - 1) // k
    for i in range(0, len(seq), step):
        sl = seq[i:i + step]
        print('%5d-%5d inst/frame %7.1f wf/frame %7.1f samples %6d  first: %s' % (i, i + len(sl), sum(s[2] for s in sl) / NF, sum(s[3] for s in sl) / NF, sum(s[5] for s in sl), sl[0][1][:60]))

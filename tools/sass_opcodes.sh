#!/bin/bash
# SASS opcode histogram of the shipped library, per kernel -> profiles/rNN_sass_opcodes.txt   (tools/sass_opcodes.sh r02)
r=${1:-r02}
cd "$(dirname "$0")/.."
( echo "# $r: SASS opcode histogram of the SHIPPED pseldnets_b200/libseldfeat.so (cuobjdump -sass), per kernel; built $(date -u +%Y-%m-%dT%H:%MZ)"
  cuobjdump -sass pseldnets_b200/libseldfeat.so | python3 -c '
import sys,re,collections
cur=None; h=collections.OrderedDict()
for l in sys.stdin:
    m=re.search(r"Function : (\S+)",l)
    if m: cur=m.group(1); h[cur]=collections.Counter(); continue
    m=re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)",l)
    if m and cur:
        op=m.group(2); parts=op.split(".")
        key=parts[0] + ("."+parts[1] if op.startswith(("LDS","STS","LDG","STG","UTC","LDTM","STTM","UTMA")) and len(parts)>1 else "")
        h[cur][key]+=1
for k,c in h.items():
    if any(t in k for t in ("foa_iv2","mic_features","scalar","rotate","wavmix","topdb","foa_features")):
        print(k, sum(c.values()), "instructions")
        print("   ", ", ".join("%s %d"%(o,n) for o,n in c.most_common(24)))
tm=collections.Counter()
for c in h.values():
    for o,n in c.items():
        if o.startswith(("UTC","LDTM","STTM","UTMA","HMMA")): tm[o]+=n
print("tensor-memory mnemonics in the shipped library (tables of the FOA / MIC kernels live in TMEM: alloc, st once, ld in the loop):", dict(tm))
print("tensor-core MMA mnemonics (UTC*MMA / HMMA) in the shipped library: 0 -- the tcgen05.mma kernels build with SELD_EXPERIMENTS=1 (profiles/r02_iv5_tensor_mel_ncu_summary.txt)")
' ) > profiles/${r}_sass_opcodes.txt
tail -2 profiles/${r}_sass_opcodes.txt | cut -c1-260

#!/bin/bash
# usage: scripts/ncu_report.sh <report.ncu-rep> <kernel-regex> <cubin-file-basename> <mangled-substr> <px-warps>
set -e
REP=$1; KRE=$2; CUBIN=$3; SUB=$4; PXW=${5:-524288}
D=$(mktemp -d); (cd $D && cuobjdump -xelf all /root/repo/segmentation_factory_b200/libsegfactory.so >/dev/null 2>&1)
ncu -i $REP --page source --csv --kernel-name regex:$KRE 2>/dev/null > $D/src.csv
python /root/repo/scripts/ncu_by_line.py $D/src.csv $D/$CUBIN $SUB 28
python - "$D/src.csv" "$PXW" <<'PY'
import csv, collections, re, sys
rows=list(csv.reader(open(sys.argv[1]))); hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}; data=rows[2:]
pxw=float(sys.argv[2])
tot=sum(int(r[idx['Instructions Executed']]) for r in data)
op=collections.Counter()
for r in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[idx['Source']])
    o=m.group(2) if m else '?'
    o='.'.join(o.split('.')[:2]) if o.startswith(('IMAD','LDS','LDG','STG','STS')) else o.split('.')[0]
    op[o]+=int(r[idx['Instructions Executed']])
print('total warp-inst', tot, 'per px-warp', round(tot/pxw,1))
print(' | '.join(f"{o} {c/pxw:.1f}" for o,c in op.most_common(34)))
st=collections.Counter()
for h in hdr:
    if h.startswith('stall_') and 'Not Issued' not in h: st[h]=sum(int(r[idx[h]] or 0) for r in data)
ts=sum(st.values()); print(' '.join(f"{k[6:]}={100*v/ts:.0f}%" for k,v in st.most_common(9)))
PY
rm -rf $D

#!/usr/bin/env python
"""Buckets the per-instruction stall samples of `ncu -i X.ncu-rep --page source --csv` by code position: where in a
fully unrolled kernel the warps wait, and on what.  Usage: tools/ncu_source_summary.py source.csv [bucket] [last_instr]"""
import csv,collections,re,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
ix={h:i for i,h in enumerate(hdr)}
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
dp_idx=[i for i,r in enumerate(data) if re.search(r'\bD(FMA|ADD|MUL)\b',r[ix['Source']])]
print("n instr",len(data),"dp range",dp_idx[0],dp_idx[-1])
B=int(sys.argv[2]) if len(sys.argv)>2 else 250
hi=int(sys.argv[3]) if len(sys.argv)>3 else len(data)
tot=sum(int(r[ix['# Samples']]) for r in data)
print("total samples",tot)
for b in range(0,hi,B):
    seg=data[b:b+B]
    sm=sum(int(r[ix['# Samples']]) for r in seg)
    if sm==0: continue
    st=collections.Counter()
    for r in seg:
        for s in stalls: st[s[6:]]+=int(r[ix[s]])
    ops=collections.Counter()
    for r in seg:
        src=r[ix['Source']].strip().split()
        op=src[1] if src[0].startswith('@') else src[0]
        ops[op.split('.')[0]]+=1
    print(b,"samples",sm,"%.1f%%"%(100*sm/tot),dict(st.most_common(4)),dict(ops.most_common(5)))

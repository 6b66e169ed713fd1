import csv, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
stall_cols=[i for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
tot=collections.Counter()
for r in data:
    for i in stall_cols:
        try: tot[hdr[i]]+=int(r[i])
        except: pass
print(tot.most_common(12))
iS=hdr.index('# Samples'); iE=hdr.index('Instructions Executed')
top=sorted(data,key=lambda r:-int(r[iS] or 0))[:int(sys.argv[2]) if len(sys.argv)>2 else 25]
for r in top:
    st={hdr[i]:int(r[i]) for i in stall_cols if r[i] not in ('','0')}
    print(r[iS], r[iE], r[1].strip()[:70], dict(sorted(st.items(), key=lambda kv:-kv[1])[:2]))
ops=collections.Counter()
for r in data:
    op=r[1].strip().split()
    if not op: continue
    o=op[1] if op[0].startswith('@') else op[0]
    ops[o]+=int(r[iE] or 0)
print([(k, round(v/524288,1)) for k,v in ops.most_common(30)])
print('total samples', sum(int(r[iS] or 0) for r in data))

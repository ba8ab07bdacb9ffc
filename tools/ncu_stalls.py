import csv, re, collections, sys
rows=list(csv.reader(open(sys.argv[1])))
hdr_idx=[i for i,r in enumerate(rows) if r and r[0]=="Address"]
h=rows[hdr_idx[0]]
end = hdr_idx[1]-1 if len(hdr_idx)>1 else len(rows)
data=rows[hdr_idx[0]+1:end]
col={n:i for i,n in enumerate(h)}
tot=collections.Counter(); byop=collections.Counter(); byop_reason=collections.defaultdict(collections.Counter); execd=collections.Counter()
stalls=[n for n in h if n.startswith('stall_') and 'Not Issued' not in n]
for r in data:
    if len(r)<len(h): continue
    src=r[col['Source']]
    m=re.match(r'\s*(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)',src)
    op=m.group(1).split('.')[0] if m else '?'
    s=int(r[col['# Samples']] or 0)
    byop[op]+=s; execd[op]+=int(r[col['Instructions Executed']] or 0)
    for n in stalls:
        v=int(r[col[n]] or 0)
        tot[n]+=v; byop_reason[op][n]+=v
T=sum(byop.values()); X=sum(execd.values())
print("total samples",T,"instructions executed",X)
for n,v in tot.most_common(10): print("%-28s %8d %.3f"%(n,v,v/T))
print()
for op,v in byop.most_common(18):
    top=", ".join("%s %.0f%%"%(k[6:],100*x/v) for k,x in byop_reason[op].most_common(4))
    print("%-10s samples %.3f  executed %.3f  %s"%(op,v/T,execd[op]/X,top))

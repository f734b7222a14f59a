import csv,sys
from collections import defaultdict
f=sys.argv[1]; fn_filter=sys.argv[2] if len(sys.argv)>2 else None
rows=list(csv.reader(open(f)))
hdr=None; cur_file=None; agg=defaultdict(lambda:[0,0,'']); section=0
for r in rows:
    if not r: continue
    if r[0]=='File Path': cur_file=r[1]; section+=1; continue
    if r[0]=='Function Name': continue
    if r[0]=='Line No': hdr=r; continue
    if hdr is None or len(r)<len(hdr)-2: continue
    if r[2]!='-': continue   # only source-line rows (sass rows have address)
    try:
        ln=int(r[0]); samples=int(r[hdr.index('# Samples')]); inst=int(r[hdr.index('Instructions Executed')])
    except: continue
    if section>1 and False: pass
    k=(cur_file.split('/')[-1],ln)
    a=agg[k]; a[0]+=samples; a[1]+=inst; a[2]=r[1]
# results duplicated per launch (2 launches) -> sum
tot=sum(a[0] for a in agg.values())
print('total samples',tot)
for k,a in sorted(agg.items(), key=lambda kv:-kv[1][0])[:int(sys.argv[3]) if len(sys.argv)>3 else 40]:
    print(f"{k[0]}:{k[1]:4d} {100*a[0]/tot:5.1f}% inst={a[1]:9d}  {a[2].strip()[:110]}")

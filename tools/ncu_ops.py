import csv, sys, subprocess
from collections import Counter
rep=sys.argv[1]; kname=sys.argv[2]; per=float(sys.argv[3])
raw=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','-k','regex:'+kname],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
# find header row
hi=[i for i,r in enumerate(rows) if r and r[0]=='Address'][0]
hdr=rows[hi]; data=[r for r in rows[hi+1:] if len(r)==len(hdr)]
ia=hdr.index("Instructions Executed"); isrc=hdr.index("Source"); ist=hdr.index("Warp Stall Sampling (All Samples)")
tot=sum(int(r[ia]) for r in data)
c=Counter(); s=Counter()
for r in data:
    t=r[isrc].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.rstrip(';')
    base=op.split('.')[0]
    if base in('LDS','STS','LDG','STG','MUFU'): base=op
    c[base]+=int(r[ia]); s[base]+=int(r[ist])
print("total", tot, "per unit", tot/per)
for op,n in c.most_common(32): print(f"{op:22s} {n/per:8.2f}  stall {s[op]}")
if len(sys.argv)>4:
    # dump hot region with stalls
    for r in data:
        if int(r[ist])>int(sys.argv[4]): print(r[ist].rjust(6), r[ia].rjust(9), r[isrc][:100])

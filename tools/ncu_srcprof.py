"""Source-level stall profile of an .ncu-rep (ncu --import-source on): hottest SASS lines with their dominant stall reasons and the
sample / executed-instruction distribution over the kernel in windows of W instructions.
   python tools/ncu_srcprof.py X.ncu-rep [ntop] [W]"""
import csv, sys, subprocess
rep = sys.argv[1]; ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=rows[1]; data=rows[2:]
iS=hdr.index("# Samples"); iI=hdr.index("Instructions Executed"); isrc=hdr.index("Source")
tot=sum(int(r[iS]) for r in data); print("total samples",tot,"n inst",len(data), "inst exec", sum(int(r[iI]) for r in data))
stall_cols=[(h,hdr.index(h)) for h in hdr if h.startswith('stall_') and 'Not' not in h]
agg={}
for r in data:
    for h,i in stall_cols:
        if r[i].isdigit(): agg[h]=agg.get(h,0)+int(r[i])
print({k:round(100*v/tot,1) for k,v in sorted(agg.items(), key=lambda x:-x[1])[:8]})
top=sorted(range(len(data)), key=lambda i:-int(data[i][iS]))[:ntop]
for i in sorted(top):
    r=data[i]
    st={h:int(r[j]) for h,j in stall_cols if r[j].isdigit() and int(r[j])>0}
    st=dict(sorted(st.items(), key=lambda x:-x[1])[:2])
    print(i, r[isrc].strip()[:64].ljust(64), r[iS], r[iI], st)
W=int(sys.argv[3]) if len(sys.argv)>3 else 100
for a in range(0,len(data),W):
    s=sum(int(r[iS]) for r in data[a:a+W]); ins=sum(int(r[iI]) for r in data[a:a+W])
    print(a, s, round(100*s/tot,1), ins)

import csv, sys
fn, kname = sys.argv[1], sys.argv[2]
import subprocess
out = subprocess.run(["ncu","-i",fn,"--page","source","--csv","--kernel-name",f"regex:{kname}"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None; data=[]; nk=0
for r in rows:
    if r and r[0]=="Kernel Name":
        nk+=1; continue
    if r and r[0]=="Address": hdr=r; continue
    if nk==1 and hdr and len(r)==len(hdr) and r[0].startswith("0x"): data.append(r)
ix={h:i for i,h in enumerate(hdr)}
recs=[(i,r[ix["Source"]].strip(),int(r[ix["Instructions Executed"]]),int(r[ix["# Samples"]]),r[ix["L1 Wavefronts Shared"]],r[ix["Avg. Threads Executed"]],r[ix["Avg. Predicated-On Threads Executed"]]) for i,r in enumerate(data)]
tot=sum(r[2] for r in recs); ts=sum(r[3] for r in recs)
print("total inst",tot,"samples",ts,"n",len(recs))
if len(sys.argv)>3:
    lo,hi=int(sys.argv[3]),int(sys.argv[4])
    for r in recs[lo:hi]: print(f"{r[0]:5d} {r[1][:58]:58s} {r[2]:9d} {r[3]:6d} wf={r[4]:>9s} thr={r[5]} on={r[6]}")
    sys.exit()
reg=[];cur=[recs[0]]
for r in recs[1:]:
    a=cur[-1][2]; b=r[2]
    if (a==0 and b==0) or (a>0 and abs(b-a)/a<0.12): cur.append(r)
    else: reg.append(cur);cur=[r]
reg.append(cur)
for g in reg:
    s=sum(r[2] for r in g); sm=sum(r[3] for r in g)
    if s/tot>0.004 or sm/ts>0.004:
        wf=sum(int(r[4]) for r in g if r[4].isdigit())
        print(f"{g[0][0]:5d}-{g[-1][0]:5d} n={len(g):3d} exec={g[0][2]:9d} inst%={100*s/tot:5.1f} samp%={100*sm/ts:5.1f} thr={g[0][5]:>5s} wf={wf:10d}  {g[0][1][:44]}")

import csv, subprocess, sys
rep=sys.argv[1]
out=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=None; fname=""; agg=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)!=len(hdr) or not r[0]: continue
    agg.append((fname,int(r[0]),r))
si=hdr.index("# Samples"); ie=hdr.index("Instructions Executed")
regions=eval(sys.argv[2])
tot_i=sum(int(r[ie] or 0) for _,_,r in agg); tot_s=sum(int(r[si] or 0) for _,_,r in agg)
acc={}
for f,ln,r in agg:
    key="other:"+f
    if f=="sweep_lean.cuh":
        key="lean:unassigned"
        for name,(a,b) in regions.items():
            if a<=ln<=b: key=name
    acc.setdefault(key,[0,0]); acc[key][0]+=int(r[ie] or 0); acc[key][1]+=int(r[si] or 0)
for k,(i,s) in sorted(acc.items(), key=lambda t:-t[1][0]):
    print("%-28s inst %5.1f%%  samples %5.1f%%" % (k, 100*i/tot_i, 100*s/tot_s))
print("total inst", tot_i)

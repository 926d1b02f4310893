import csv, subprocess, sys
rep=sys.argv[1]; topn=int(sys.argv[2]) if len(sys.argv)>2 else 40
out=subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines())); hdr=None; fname=""; agg=[]
for r in rows:
    if len(r)==2 and r[0]=="File Path": fname=r[1].split("/")[-1]; continue
    if len(r)>5 and r[0]=="Line No": hdr=r; continue
    if hdr is None or len(r)!=len(hdr) or not r[0]: continue
    agg.append((fname,r))
si=hdr.index("# Samples"); ie=hdr.index("Instructions Executed"); it=hdr.index("Thread Instructions Executed") if "Thread Instructions Executed" in hdr else None
toti=sum(int(r[ie] or 0) for _,r in agg); tot=sum(int(r[si] or 0) for _,r in agg)
print("total inst", toti)
for f,r in sorted(agg,key=lambda t:-int(t[1][ie] or 0))[:topn]:
    ti = int(r[it] or 0)/max(1,int(r[ie] or 0)) if it else 0
    print("%5.1f%% inst %5.1f%% smp lanes %4.1f %s:%s | %s" % (100*int(r[ie] or 0)/toti, 100*int(r[si] or 0)/tot, ti, f, r[0], r[1].strip()[:110]))

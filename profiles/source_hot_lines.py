import csv,sys,subprocess
rep=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 50
out=subprocess.run(["ncu","-i",rep,"--page","source","--csv","--print-source","cuda,sass"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; agg={}; hdr=None
for r in rows:
    if len(r)==2 and r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r and r[0]=="Line No": hdr=r; continue
    if hdr and len(r)==len(hdr) and cur:
        try: s=int(r[hdr.index("# Samples")]); ln=int(r[0])
        except: continue
        ie=int(r[hdr.index("Instructions Executed")] or 0)
        wfx=int(r[hdr.index("L1 Wavefronts Shared Excessive")] or 0); wf=int(r[hdr.index("L1 Wavefronts Shared")] or 0)
        a=agg.setdefault((cur,ln),[0,0,r[1][:100],0,0])
        a[0]+=s;a[1]+=ie;a[3]+=wfx;a[4]+=wf
tot=sum(v[0] for v in agg.values()); ti=sum(v[1] for v in agg.values())
print("total samples",tot,"total warp-instr",ti)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(f"{k[0][:14]}:{k[1]:4d} {100*v[0]/tot:5.1f}% inst={v[1]:8d} wfx={v[3]:8d} wf={v[4]:8d} | {v[2]}")

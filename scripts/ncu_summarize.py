import csv,sys,collections,re,subprocess
rep=sys.argv[1]
raw=subprocess.run(["ncu","-i",rep,"--page","raw","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
h=rows[0]; v=rows[2]
out=[]
for i,k in enumerate(h):
    if k in ("gpu__time_duration.sum","smsp__issue_active.avg.per_cycle_active","smsp__average_warp_latency_per_inst_issued.ratio","smsp__inst_executed.sum","launch__registers_per_thread","sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active") or ('issue_stalled' in k and 'per_issue_active' in k and float(v[i] or 0)>0.2):
        out.append("%s=%s"%(k.replace('smsp__average_warps_issue_stalled_','').replace('_per_issue_active.ratio',''),v[i][:8]))
print(" ".join(out))
src=subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(src.splitlines()))
h=rows[1]
ia=h.index('Address'); isrc=h.index('Source'); ino=h.index('stall_no_inst'); iex=h.index('Instructions Executed'); ins=h.index('# Samples')
data=[]
for r in rows[2:]:
    if len(r)<=ino: continue
    data.append((int(r[ia],16), r[isrc].strip(), int(r[ino] or 0), int(r[iex] or 0), int(r[ins] or 0)))
base=data[0][0]
c=collections.Counter()
for a,s,n,e,sm in data: c[((a-base)%128)//16]+=n
nw=data[0][3]
print("samples",sum(d[4] for d in data),"no_inst",sum(d[2] for d in data),"by pos",[c[i] for i in range(8)],"instr/warp %.0f"%(sum(d[3] for d in data)/nw))
# lines touched (executed>0) and their utilization
lines=collections.defaultdict(list)
for a,s,n,e,sm in data: lines[(a-base)//128].append(e)
touched=[l for l in lines.values() if max(l)>0]
print("static lines",len(lines),"touched lines",len(touched),"KB touched %.0f"%(len(touched)*128/1024),"avg util %.2f"%(sum(sum(1 for e in l if e>0.5*nw) for l in touched)/ (8*len(touched))))
oc=collections.Counter()
for a,s,n,e,sm in data:
    s=re.sub(r'^@!?U?P\d+\s+','',s); oc[s.split()[0].split('.')[0]]+=e
print(" ".join("%s=%.0f"%(k,v/nw) for k,v in oc.most_common(16)))

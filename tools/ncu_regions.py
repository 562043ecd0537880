"""Stall samples of an `ncu --set full --import-source on` capture aggregated per region between barriers / mbarrier waits:
`ncu -i x.ncu-rep --page source --csv --print-source sass > x.csv; python tools/ncu_regions.py x.csv` (share of all samples, samples on
the first instruction of the region = the wait itself, stall reasons above 5 %)."""
import csv,sys
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]; data=rows[2:]
isrc=hdr.index("Source"); isamp=hdr.index("Warp Stall Sampling (All Samples)"); iex=hdr.index("Instructions Executed")
cols=['stall_barrier','stall_branch_resolving','stall_dispatch','stall_drain','stall_lg','stall_long_sb','stall_math','stall_membar','stall_mio','stall_misc','stall_no_inst','stall_not_selected','stall_selected','stall_short_sb','stall_sleep','stall_tex','stall_wait']
idx={c:hdr.index(c) for c in cols}
tot=sum(int(r[isamp]) for r in data)
print("total samples",tot,"n instr",len(data))
acc=0; start=0; regions=[]
for i,r in enumerate(data):
    s=r[isrc]; acc+=int(r[isamp])
    if "BAR.SYNC" in s or "SYNCS.PHASECHK" in s:
        regions.append((start,i,acc,s.strip()[:50])); start=i+1; acc=0
regions.append((start,len(data)-1,acc,"end"))
def agg(a,b):
    t={c:0 for c in cols}
    for r in data[a:b+1]:
        for c in cols: t[c]+=int(r[idx[c]] or 0)
    tt=sum(t.values())
    return {c.replace('stall_',''):round(100*v/max(tt,1)) for c,v in t.items() if v>tt*0.05}
for a,b,c,s in regions:
    if c>tot*0.004: print(f"{a:5d}-{b:5d} {c:7d} {100*c/tot:5.1f}%  first-instr {int(data[a][isamp]):6d} {s[:30]:30s} {agg(a,b)}")

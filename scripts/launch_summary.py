"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) per kernel."""
import csv, collections, re, sys
for f in sys.argv[1:]:
    rows=[r for r in csv.reader(open(f)) if len(r)>10 and r[0].isdigit()]
    agg=collections.OrderedDict()
    for r in rows:
        name=re.sub(r"\(.*","",r[4])
        m=re.search(r"(wave_fwd_kernel|wave_bwd_kernel|small_kernel|small_hom_kernel|wsc_kernel|wsc_block_kernel)<([\d, ]+)>", r[4])
        if m: name=f"{m.group(1)}<{m.group(2)}>"
        key=(name,r[7],)
        a=agg.setdefault(key,[0,0.0,0])
        a[0]+=1; a[1]+=float(r[-1])/1e6; a[2]=max(a[2],int(r[8].strip("()").split(",")[0]))
    tot=sum(a[1] for a in agg.values())
    print(f, "n launches", len(rows), "total ms", round(tot,2))
    for k,a in agg.items(): print(f"  {k[0][:60]:60s} blk{k[1]:14s} n={a[0]:4d} ms={a[1]:9.3f} share={a[1]/tot*100:5.1f}% maxgrid={a[2]}")

out=gpurun_out/$1; mkdir -p $out
VD_WSC_SPLIT=1 python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 10 > $out/bench_split1.json 2> $out/bench_split1.err
python -c "
import json; d=json.load(open('$out/bench_split1.json')); print('split=1 step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['serial_pass_ms'], d['gpu_launches'])"
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
VD_SERIAL=1 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $out/launches_split.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary --no-seam > $out/launches_split.log 2>&1
CSV=$out/launches_split.csv python - <<'P'
import csv, re, collections, sys, os
rows=[r for r in csv.reader(open(os.environ['CSV'])) if len(r)>10 and r[0].isdigit()]
agg=collections.OrderedDict(); steps=0
for r in rows:
    name=re.sub(r"\(vd::.*","",r[4]); name=re.sub(r"\(int\)|\(bool\)","",name)
    if 'plan_kernel' in name: steps+=1
    a=agg.setdefault(name,[0,0.0]); a[0]+=1; a[1]+=float(r[-1])/1e6
for k,a in agg.items():
    if 'at::' in k: continue
    print(f"{k[:70]:70s} n={a[0]:3d} ms/step={a[1]/max(steps,1):7.3f}")
print("steps", steps)
P

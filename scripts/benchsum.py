import sys, json
for l in sys.stdin:
    if not l.startswith("{"): continue
    d = json.loads(l)
    r = d["roofline"]
    print(f"value {d['value']:.2f} Gcells/s ({d['ms_per_step']:.2f} ms)  e2e {d['e2e']['value']:.2f} ({d['e2e']['ms_per_step']:.2f} ms)  "
          f"serial {r.get('serial_step_ms', 0):.2f} ms {json.dumps({k: round(v, 2) for k, v in r.get('serial_pass_ms', {}).items()})}  "
          f"dom {r['kernel']} {r['achieved']:.1f} GB/s frac {r['frac']:.4f}")

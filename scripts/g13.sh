out=gpurun_out/$1; mkdir -p $out
for v in 1 0; do
VD_PRIO=$v python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 20 > $out/bench_prio$v.json 2> $out/bench_prio$v.err; tail -2 $out/bench_prio$v.err
python -c "
import json; d=json.load(open('$out/bench_prio$v.json')); print('prio=$v step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['clocks'])"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -2

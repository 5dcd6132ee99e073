out=gpurun_out/$1; mkdir -p $out
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 10 > $out/bench_quick.json 2> $out/bench_quick.err; tail -3 $out/bench_quick.err
python -c "
import json; d=json.load(open('$out/bench_quick.json')); print('step', round(d['ms_per_step'],3)); print(json.dumps(d['e2e'], indent=1))"

# wgs: split vs fused warp path, parity
out=gpurun_out/$1; mkdir -p $out
for v in 1 0; do
VD_WSC_SPLIT=$v python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 10 > $out/bench_split$v.json 2> $out/bench_split$v.err
python -c "
import json; d=json.load(open('$out/bench_split$v.json')); print('split=$v step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['serial_pass_ms'], d['gpu_launches'])"
done
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3

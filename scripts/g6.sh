# wgs: quick bench + phase split + parity
out=gpurun_out/$1; mkdir -p $out
python bench.py --no-cpu-baseline --no-secondary --no-seam --steps 10 > $out/bench_quick.json 2> $out/bench_quick.err
python -c "
import json; d=json.load(open('$out/bench_quick.json')); print('step', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['ms_per_step'],2), d['roofline']['serial_pass_ms'])"
VD_LIB=vcfdist_b200/libvd_prof.so python scripts/phases.py 3600000 > $out/phases.txt 2>&1; cat $out/phases.txt
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3

mkdir -p gpurun_out/e15
{
VD_LIB=vcfdist_b200/libvd_e1.so timeout 1500 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "long_alignments or sv_lengths or divergent or dense_paths or wgs_like or golden_through" 2>&1 | tail -12
} > gpurun_out/e15/log 2>&1; cat gpurun_out/e15/log

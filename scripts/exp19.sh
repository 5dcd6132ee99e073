mkdir -p gpurun_out/e19
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 3 --warmup 3 2>&1 | grep '^{' | tee gpurun_out/e19/bench_8gpu.json | python scripts/benchsum.py
} > gpurun_out/e19/log 2>&1; tail -c 2000 gpurun_out/e19/log

mkdir -p gpurun_out/e10
{
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 4 --steps 3 --warmup 3 --gather-slices 1 2>&1 | grep '^{' | python scripts/benchsum.py
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 3 --warmup 3 2>&1 | grep '^{' | python scripts/benchsum.py
} > gpurun_out/e10/log4 2>&1; tail -c 3000 gpurun_out/e10/log4

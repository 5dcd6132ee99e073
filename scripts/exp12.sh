mkdir -p gpurun_out/e12
{
python scripts/exp6.py
VD_RAMP=0 python scripts/exp6.py
VD_RAMP=0 VD_CHUNK_SC=1300000 python scripts/exp6.py
VD_CHUNK_SC=1500000 python scripts/exp6.py
VD_RAMP=0 VD_CHUNK_SC=2000000 python scripts/exp6.py
VD_RAMP=0 VD_CHUNK_SC=700000 python scripts/exp6.py
} 2>&1 | grep -v run_resident > gpurun_out/e12/log; cat gpurun_out/e12/log

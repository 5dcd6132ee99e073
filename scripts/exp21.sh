mkdir -p gpurun_out/e21
{
timeout 200 python -m pytest tests -m gpu -q -x -k "golden_through_c_abi or edge_cases or device_slices or cli_dropin" 2>&1 | tail -3
timeout 100 python -m pytest tests -m "not gpu" -q -x -k "known_answers or exports or product_package" 2>&1 | tail -2
} > gpurun_out/e21/log 2>&1; cat gpurun_out/e21/log

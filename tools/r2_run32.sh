timeout 900 python -m pytest tests/test_dropin.py tests/test_integration_patch.py tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_t32.log 2>&1; tail -4 gpurun_out/r2_t32.log

python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" > gpurun_out/r2_t3.log 2>&1; tail -3 gpurun_out/r2_t3.log
bash tools/r2_ab.sh

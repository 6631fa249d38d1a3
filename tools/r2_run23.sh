timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "parity_vs_oracle or knobs or extreme or odd" > gpurun_out/r2_t23.log 2>&1; tail -2 gpurun_out/r2_t23.log
bash tools/r2_quick.sh

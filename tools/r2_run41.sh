timeout 1500 python tools/fuzz_parity.py --seeds 400 --first 20000 > gpurun_out/r2s3_fuzz_long.log 2>&1; tail -2 gpurun_out/r2s3_fuzz_long.log

# final kernel of round 2: full GPU suite, a long fuzz campaign, the default bench line
set -x
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2s3_gputests_final.log 2>&1; tail -4 gpurun_out/r2s3_gputests_final.log
timeout 1200 python tools/fuzz_parity.py --seeds 150 --first 9000 > gpurun_out/r2s3_fuzz_final.log 2>&1; tail -2 gpurun_out/r2s3_fuzz_final.log
( time python bench.py ) > gpurun_out/r2s3_bench_final_n1.json 2> gpurun_out/r2s3_bench_final_n1.err; tail -3 gpurun_out/r2s3_bench_final_n1.err
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r2s3_bench_ref_n1.json 2> gpurun_out/r2s3_bench_ref_n1.err; tail -2 gpurun_out/r2s3_bench_ref_n1.err

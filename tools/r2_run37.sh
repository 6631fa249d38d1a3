# end of round 2: smoke, full GPU suite, default bench + reference arm, launch list of the bench command
set -x
python __graft_entry__.py --smoke > gpurun_out/r2s3_smoke.log 2>&1; tail -2 gpurun_out/r2s3_smoke.log
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2s3_gputests_final.log 2>&1; tail -4 gpurun_out/r2s3_gputests_final.log
( time python bench.py ) > gpurun_out/r2s3_bench_final_n1.json 2> gpurun_out/r2s3_bench_final_n1.err; tail -3 gpurun_out/r2s3_bench_final_n1.err
python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_bench_plain37.json 2> gpurun_out/r2_bench_plain37.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2s3_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_ncu_bench37.log 2>&1

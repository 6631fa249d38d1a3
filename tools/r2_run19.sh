# round 2, session 3: slot-walk kernel -- full GPU suite, fuzz, default bench, launch list, ncu captures
set -x
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_gputests19.log 2>&1; tail -4 gpurun_out/r2_gputests19.log
timeout 900 python tools/fuzz_parity.py --seeds 40 --first 8000 > gpurun_out/r2_fuzz19.log 2>&1; tail -2 gpurun_out/r2_fuzz19.log
( time python bench.py ) > gpurun_out/r2_bench19.json 2> gpurun_out/r2_bench19.err; tail -3 gpurun_out/r2_bench19.err
python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_bench_plain19.json 2> gpurun_out/r2_bench_plain19.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches19.csv python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_ncu_bench19.log 2>&1
bash tools/prof.sh s3city --map city10k --kind long --n 16777216
bash tools/prof.sh s3rooms --map rooms8 --kind uniform --n 16777216
bash tools/prof.sh s3cityshort --map city10k --kind short --n 16777216
bash tools/prof.sh s3mega --map mega200k --kind short --n 16777216

# round 2, session 2: full GPU suite, full default bench (timed), ncu captures of the current kernel
set -x
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2_gputests.log 2>&1; tail -8 gpurun_out/r2_gputests.log
( time python bench.py ) > gpurun_out/r2_bench_full.json 2> gpurun_out/r2_bench_full.err; tail -5 gpurun_out/r2_bench_full.err
bash tools/prof.sh r2city --map city10k --kind long --n 16777216
bash tools/prof.sh r2rooms --map rooms8 --kind uniform --n 16777216

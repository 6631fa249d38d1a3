set -x
python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_bench_plain.json 2> gpurun_out/r2_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_city10k.csv python bench.py --steps 2 --warmup 3 --no-cpu --also "" > gpurun_out/r2_ncu_bench.log 2>&1
tail -2 gpurun_out/r2_bench_plain.err
bash tools/prof.sh r2mega --map mega200k --kind short --n 16777216
python bench.py > gpurun_out/r2_bench_full2.json 2> gpurun_out/r2_bench_full2.err; tail -3 gpurun_out/r2_bench_full2.err

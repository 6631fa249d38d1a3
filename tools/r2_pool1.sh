set -x
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "queue_scheduler" > gpurun_out/pool_t1.log 2>&1; tail -15 gpurun_out/pool_t1.log
rm -f gpurun_out/pool_tune1.log
for spec in "city10k long" "rooms8 uniform" "mega200k short" "city10k short"; do set -- $spec; timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --sched 0,5 >> gpurun_out/pool_tune1.log 2>&1; done
cut -c1-60,200-400 gpurun_out/pool_tune1.log

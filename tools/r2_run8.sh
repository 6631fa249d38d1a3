timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size and not queue" > gpurun_out/r2_t8.log 2>&1; tail -4 gpurun_out/r2_t8.log
rm -f gpurun_out/r2_tune8.log
for spec in "city10k long" "rooms8 uniform" "mega200k short" "city10k short"; do set -- $spec; timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 >> gpurun_out/r2_tune8.log 2>&1; done
cut -c1-40,100-130,300-400 gpurun_out/r2_tune8.log

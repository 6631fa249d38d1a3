set -x
python -m pytest tests/test_gpu_parity.py -x -q -k "not full_size" > gpurun_out/r2_t2.log 2>&1; tail -5 gpurun_out/r2_t2.log
rm -f gpurun_out/r2_tune2.log
for spec in "city10k long" "rooms8 uniform" "mega200k short" "city10k short"; do set -- $spec; python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 >> gpurun_out/r2_tune2.log 2>&1; done
python tools/tune.py --map rooms8 --kind uniform --n 33554432 --reps 3 --slots 4,6 >> gpurun_out/r2_tune2.log 2>&1
python tools/tune.py --map city10k --kind long --n 33554432 --reps 3 --general 1 >> gpurun_out/r2_tune2.log 2>&1
python tools/tune.py --map city10k --kind long --n 33554432 --reps 3 --bin 1 >> gpurun_out/r2_tune2.log 2>&1
python tools/tune.py --map city10k --kind long --n 33554432 --reps 3 --slots 4,8 --steps 6,12,16 >> gpurun_out/r2_tune2.log 2>&1
python tools/tune.py --map mega200k --kind short --n 33554432 --reps 3 --bin 1 >> gpurun_out/r2_tune2.log 2>&1
cat gpurun_out/r2_tune2.log

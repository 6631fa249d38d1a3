rm -f gpurun_out/r2_tune7.log
python tools/tune.py --map mega200k --kind short --n 33554432 --reps 3 --refill 4,8,12,16,24 --steps 4,8 >> gpurun_out/r2_tune7.log 2>&1
python tools/tune.py --map city10k --kind long --n 33554432 --reps 3 --refill 6,8,12 --leaf_thr 12,16,20 >> gpurun_out/r2_tune7.log 2>&1
python tools/tune.py --map rooms8 --kind uniform --n 33554432 --reps 3 --refill 6,8,12 --leaf_thr 12,16,20 >> gpurun_out/r2_tune7.log 2>&1
cut -c1-40,150-330 gpurun_out/r2_tune7.log

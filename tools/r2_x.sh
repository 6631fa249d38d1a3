for mode in 0 1 2 3; do
  BSPGPU_TMP_X=$mode python tools/tune.py --map city10k --kind short --n 33554432 --reps 3 2>&1 | tail -1 | sed "s/^/mode$mode /"
done
python tools/tune.py --map city10k --kind short --n 33554432 --reps 1 > /dev/null 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 14 --csv --log-file gpurun_out/r2_x_launches.csv python tools/tune.py --map city10k --kind short --n 33554432 --reps 1 > /dev/null 2>&1
cat gpurun_out/r2_x_launches.csv | tail -16

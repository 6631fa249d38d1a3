rm -f gpurun_out/r2_slots2.log
for v in 0 1; do
  if [ $v = 1 ]; then export BSPGPU_SLOTS=1; else unset BSPGPU_SLOTS; fi
  python tools/tune.py --map rooms8 --kind uniform --n 33554432 --reps 3 --variants 0,3 2>&1 | tail -2 | sed "s/^/slots$v /" >> gpurun_out/r2_slots2.log
done
cut -c1-60,80-100,250-400 gpurun_out/r2_slots2.log

# child-slot walk experiment: BSPGPU_SLOTS=1 selects trace_slots for the GLOBAL variant
rm -f gpurun_out/r2_slots.log
BSPGPU_SLOTS=1 timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "parity_vs_oracle or mega200k or extreme or start_grid or odd_map" > gpurun_out/r2_slots_tests.log 2>&1; tail -3 gpurun_out/r2_slots_tests.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec
for v in 0 1; do
  if [ $v = 1 ]; then export BSPGPU_SLOTS=1; else unset BSPGPU_SLOTS; fi
  python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 2>&1 | tail -1 | sed "s/^/slots$v /" >> gpurun_out/r2_slots.log
done; done
cut -c1-60,250-400 gpurun_out/r2_slots.log

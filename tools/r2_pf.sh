# prefetch experiment: children line prefetched as soon as the node record arrives (BSPGPU_PF: 0 off, 1 L1, 2 dummy load, 3 L2)
rm -f gpurun_out/r2_pf.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec
for pf in 0 1 2 3; do
  BSPGPU_PF=$pf python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 2>&1 | tail -1 | sed "s/^/pf$pf /" >> gpurun_out/r2_pf.log
done; done
cat gpurun_out/r2_pf.log

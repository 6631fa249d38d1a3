rm -f gpurun_out/r2_pf2.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec
for pf in 0 4 8 16; do
  BSPGPU_PF=$pf python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 2>&1 | tail -1 | sed "s/^/pf$pf /" >> gpurun_out/r2_pf2.log
done; done
cut -c1-50,250-400 gpurun_out/r2_pf2.log

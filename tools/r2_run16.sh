rm -f gpurun_out/r2_tune16.log
for lr in 1000 6 5 4 3 2; do
for spec in "city10k long 10" "mega200k short 8" "city10k short 8" "mega200k long 10"; do set -- $spec
  BSPGPU_LR=$lr timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --steps $3 2>&1 | tail -1 | sed "s/^/lr$lr /" >> gpurun_out/r2_tune16.log
done; done
cut -c1-50,250-400 gpurun_out/r2_tune16.log

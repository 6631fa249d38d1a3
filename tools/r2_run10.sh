rm -f gpurun_out/r2_tune10.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec; timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 >> gpurun_out/r2_tune10.log 2>&1; done

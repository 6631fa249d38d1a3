export BSPGPU_SLOTS=1
for spec in "city10k long cityslots" "mega200k short megaslots"; do set -- $spec
python tools/tune.py --map $1 --kind $2 --n 16777216 --reps 1 > gpurun_out/plain_$3.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:trace_s -s 1 -c 1 -f -o gpurun_out/prof_$3 python tools/tune.py --map $1 --kind $2 --n 16777216 --reps 1 > gpurun_out/ncu_$3.log 2>&1
tail -n 1 gpurun_out/plain_$3.log | cut -c1-50,240-400
done

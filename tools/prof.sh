#!/bin/bash
# usage: prof.sh <name> <tune args...>
name=$1; shift
python tools/tune.py "$@" --reps 1 > gpurun_out/plain_$name.log 2>&1 && ncu --set full --clock-control none --import-source on -k "regex:trace_(p|s)" -s 1 -c 1 -f -o gpurun_out/prof_$name python tools/tune.py "$@" --reps 1 > gpurun_out/ncu_$name.log 2>&1
tail -n 1 gpurun_out/plain_$name.log

export BSPGPU_GT=4
bash tools/prof.sh boxrooms --map rooms8 --kind uniform --n 16777216
bash tools/prof.sh boxcity --map city10k --kind long --n 16777216

bash tools/prof.sh r2short --map city10k --kind short --n 16777216
cd .r1ref && mkdir -p gpurun_out && bash tools/prof.sh r1short --map city10k --kind short --n 16777216 --threads 0 --steps 0 --side_steps 3 --leaf_thr 16; cp gpurun_out/*r1short* ../gpurun_out/

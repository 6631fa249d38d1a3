# box step (six axial brush sides per leaf-step turn): parity first, then the quick workloads
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_t25.log 2>&1; tail -3 gpurun_out/r2_t25.log
bash tools/r2_quick.sh

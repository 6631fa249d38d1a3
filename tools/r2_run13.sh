timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "pageable or host_pipeline or small_batches or mixed_pointer" > gpurun_out/r2_t13.log 2>&1; tail -5 gpurun_out/r2_t13.log
python bench.py --steps 3 --warmup 3 --also "" --no-cpu > gpurun_out/r2_bench13.json 2> gpurun_out/r2_bench13.err; tail -3 gpurun_out/r2_bench13.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench13.json').read().strip().splitlines()[-1]); print(d['e2e']['value'], d['e2e']['e2e_pageable'])"
nproc

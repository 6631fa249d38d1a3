timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "compact or pageable or small_batches or host_pipeline" > gpurun_out/r2_t31.log 2>&1; tail -3 gpurun_out/r2_t31.log
python bench.py --steps 3 --warmup 3 --also "" --no-cpu > gpurun_out/r2_bench31.json 2> gpurun_out/r2_bench31.err; tail -2 gpurun_out/r2_bench31.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench31.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['layout'], {k:v['value'] for k,v in d['e2e']['layouts'].items()})"

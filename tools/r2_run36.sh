timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_dropin.py -x -q -m gpu -k "small_batches or compact or dropin or pageable or concurrent" > gpurun_out/r2_t36.log 2>&1; tail -5 gpurun_out/r2_t36.log
python bench.py --steps 3 --warmup 3 --also "" --no-cpu > gpurun_out/r2_bench36.json 2> gpurun_out/r2_bench36.err; tail -2 gpurun_out/r2_bench36.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench36.json').read().strip().splitlines()[-1]); print(d['e2e']['latency'])"

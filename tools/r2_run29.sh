# memcheck of the slot-walk kernel on small maps (GLOBAL and SMEM variants), then the default bench with the shuffled camera order
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -x -q -k "parity_vs_oracle and (kat1 or tilted)" > gpurun_out/r2s3_memcheck.log 2>&1; echo "memcheck rc=$?"; tail -6 gpurun_out/r2s3_memcheck.log
( time python bench.py ) > gpurun_out/r2s3_bench29.json 2> gpurun_out/r2s3_bench29.err; tail -3 gpurun_out/r2s3_bench29.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s3_bench29.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], {k:(v.get('value'), v.get('parity_sample',{}).get('mismatches'), v.get('error')) for k,v in d['also'].items()})
PY

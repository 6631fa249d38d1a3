( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29524 bench.py --gpus 4 ) > gpurun_out/r2s3_bench_n4.json 2> gpurun_out/r2s3_bench_n4.err; tail -3 gpurun_out/r2s3_bench_n4.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s3_bench_n4.json').read().strip().splitlines()[-1]); print(d['value'], d['parity_sample']['checked'], d['parity_sample']['mismatches'], d['e2e']['layout'], {k:v['value'] for k,v in d['e2e']['layouts'].items()}, {k:v['value'] for k,v in d['also'].items()})"

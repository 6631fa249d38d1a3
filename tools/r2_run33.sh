# 8 GPUs: e2e with the 4-byte result layout
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 --also "" ) > gpurun_out/r2s3_bench_n8b.json 2> gpurun_out/r2s3_bench_n8b.err; tail -4 gpurun_out/r2s3_bench_n8b.err
python -c "
import json; d=json.loads(open('gpurun_out/r2s3_bench_n8b.json').read().strip().splitlines()[-1]); print(d['value'], d['e2e']['layout'], {k:v['value'] for k,v in d['e2e']['layouts'].items()})"

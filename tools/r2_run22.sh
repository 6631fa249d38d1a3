# 8 GPUs, session 3 (slot-walk kernel)
set -x
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 ) > gpurun_out/r2s3_bench_n8.json 2> gpurun_out/r2s3_bench_n8.err; tail -4 gpurun_out/r2s3_bench_n8.err
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2s3_multi_n8.log 2>&1; tail -3 gpurun_out/r2s3_multi_n8.log

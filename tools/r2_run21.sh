# 2 GPUs, session 3 (slot-walk kernel): multi-GPU entry points + bench under torchrun
set -x
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2s3_multi_n2.log 2>&1; tail -3 gpurun_out/r2s3_multi_n2.log
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 ) > gpurun_out/r2s3_bench_n2.json 2> gpurun_out/r2s3_bench_n2.err; tail -4 gpurun_out/r2s3_bench_n2.err
head -c 600 gpurun_out/r2s3_bench_n2.json

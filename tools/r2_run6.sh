# 2 GPUs: the multi-GPU entry points + bench under torchrun (parity sample gathered over the library's NCCL)
set -x
nvidia-smi -L
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi_n2.log 2>&1; tail -5 gpurun_out/r2_multi_n2.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n2.json 2> gpurun_out/r2_bench_ref_n2.err; tail -3 gpurun_out/r2_bench_ref_n2.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 ) > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; tail -5 gpurun_out/r2_bench_n2.err
head -c 1500 gpurun_out/r2_bench_n2.json

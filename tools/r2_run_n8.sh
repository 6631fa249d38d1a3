# 8 GPUs: topology / link ceiling, the reference arm, our arm (default line: every N-rank parity sample, C5 included)
set -x
bash tools/topo_probe.sh > gpurun_out/r2_topo_n8.log 2>&1; tail -3 gpurun_out/r2_topo_n8.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref_n8.json 2> gpurun_out/r2_bench_ref_n8.err; tail -2 gpurun_out/r2_bench_ref_n8.err
( time python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 8 ) > gpurun_out/r2_bench_n8.json 2> gpurun_out/r2_bench_n8.err; tail -5 gpurun_out/r2_bench_n8.err
python -m pytest tests/test_gpu_multi.py -x -q -m gpu > gpurun_out/r2_multi_n8.log 2>&1; tail -3 gpurun_out/r2_multi_n8.log

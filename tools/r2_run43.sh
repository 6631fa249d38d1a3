python __graft_entry__.py --smoke 2>&1 | tail -1
( time python -m pytest tests -x -q -m gpu ) > gpurun_out/r2s3_gputests_final.log 2>&1; tail -4 gpurun_out/r2s3_gputests_final.log
bash tools/r2_quick.sh

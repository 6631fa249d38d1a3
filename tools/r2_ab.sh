# same-box A/B: round-1 library (.r1ref, commit b24255d) against the working tree, device-resident, 32 Mi segments
rm -f gpurun_out/r2_ab.log
for rep in 1 2; do
for spec in "city10k long" "rooms8 uniform" "mega200k short" "city10k short"; do set -- $spec
  (cd .r1ref && python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --threads 0 --steps 0 --side_steps 3 --leaf_thr 16 2>&1 | tail -1 | sed 's/^/r1 /') >> gpurun_out/r2_ab.log
  python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 2>&1 | tail -1 | sed 's/^/r2 /' >> gpurun_out/r2_ab.log
done; done
cat gpurun_out/r2_ab.log

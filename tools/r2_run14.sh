# slot walk as the GLOBAL-variant kernel: full GPU suite, a short fuzz campaign, knob sweep
(time timeout 1500 python -m pytest tests -m gpu -x -q) > gpurun_out/r2_gputests14.log 2>&1; tail -4 gpurun_out/r2_gputests14.log
timeout 600 python tools/fuzz_parity.py --seeds 12 --first 7000 > gpurun_out/r2_fuzz14.log 2>&1; tail -2 gpurun_out/r2_fuzz14.log
rm -f gpurun_out/r2_tune14.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec; timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --slots 4,6,8 --refill 6,8,12 >> gpurun_out/r2_tune14.log 2>&1; done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune14.log'):
    try: d=json.loads(ln)
    except: continue
    print(d['map'],d['kind'],'slots',d['slots'],'refill',d['refill'],d['ms'],d['Mtraces_s'])
PY

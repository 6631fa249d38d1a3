rm -f gpurun_out/r2_tune20.log
timeout 600 python tools/tune.py --map city10k --kind camera --n 33554432 --reps 3 --steps 0,6,8,10,12,16 --leaf_thr 12,16,20 >> gpurun_out/r2_tune20.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune20.log'):
    try: d=json.loads(ln)
    except: print(ln[:200]); continue
    print(d['map'],d['kind'],'steps',d['max_steps'],'leaf_thr',d['leaf_thr'],d['ms'],d['Mtraces_s'])
PY

rm -f gpurun_out/r2_tune15.log
timeout 600 python tools/tune.py --map city10k --kind long --n 33554432 --reps 3 --steps 6,8,10,12 --leaf_thr 12,16,20 --side_steps 2,3,4 >> gpurun_out/r2_tune15.log 2>&1
timeout 300 python tools/tune.py --map mega200k --kind short --n 33554432 --reps 3 --steps 4,6,8,12 --leaf_thr 12,16 --side_steps 3 >> gpurun_out/r2_tune15.log 2>&1
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune15.log'):
    try: d=json.loads(ln)
    except: continue
    print(d['map'],d['kind'],'steps',d['max_steps'],'leaf_thr',d['leaf_thr'],'side',d['side_steps'],d['ms'],d['Mtraces_s'])
PY

rm -f gpurun_out/r2_tune18.log
for lr in 1000 4 3 2 1; do
  BSPGPU_LR=$lr timeout 300 python tools/tune.py --map rooms8 --kind uniform --n 33554432 --reps 3 --steps 8,12,16 2>&1 | tail -3 | sed "s/^/lr$lr /" >> gpurun_out/r2_tune18.log
  BSPGPU_LR=$lr timeout 300 python tools/tune.py --map rooms8 --kind short --n 33554432 --reps 3 --steps 8,16 2>&1 | tail -2 | sed "s/^/lr$lr /" >> gpurun_out/r2_tune18.log
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune18.log'):
    tag, js = ln.split(' ',1)
    try: d=json.loads(js)
    except: continue
    print(tag, d['map'],d['kind'],'steps',d['max_steps'],'leaf_thr',d['leaf_thr'],d['ms'],d['Mtraces_s'])
PY

rm -f gpurun_out/r2_tune17.log
for lr in 2 1; do
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec
  BSPGPU_LR=$lr timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --steps 8,10 --leaf_thr 12,16 2>&1 | tail -4 | sed "s/^/lr$lr /" >> gpurun_out/r2_tune17.log
done; done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune17.log'):
    tag, js = ln.split(' ',1)
    try: d=json.loads(js)
    except: continue
    print(tag, d['map'],d['kind'],'steps',d['max_steps'],'leaf_thr',d['leaf_thr'],d['ms'],d['Mtraces_s'])
PY

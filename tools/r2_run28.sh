rm -f gpurun_out/r2_tune28.log
for gt in 1 2 3; do for bs in 1 2 3; do
for spec in "city10k long" "rooms8 uniform"; do set -- $spec
  BSPGPU_GT=$gt BSPGPU_BS=$bs python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 2 --leaf_thr 12,16,20 2>&1 | tail -3 | sed "s/^/gt$gt bs$bs /" >> gpurun_out/r2_tune28.log
done; done; done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune28.log'):
    a,b,js=ln.split(' ',2)
    try: d=json.loads(js)
    except: print(ln[:200]); continue
    print(a,b,d['map'],d['kind'],'thr',d['leaf_thr'],d['ms'],d['Mtraces_s'])
PY

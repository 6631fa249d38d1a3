rm -f gpurun_out/r2_tune30.log
timeout 600 python tools/tune.py --map rooms8 --kind uniform --n 33554432 --reps 3 --steps 16,20,24,32,48 --leaf_thr 12,16,20 --side_steps 3,4 >> gpurun_out/r2_tune30.log 2>&1
python - <<'PY'
import json
rows=[]
for ln in open('gpurun_out/r2_tune30.log'):
    try: d=json.loads(ln)
    except: continue
    rows.append((d['Mtraces_s'],d['max_steps'],d['leaf_thr'],d['side_steps']))
for r in sorted(rows)[-12:]: print(r)
print(sorted(rows)[0])
PY

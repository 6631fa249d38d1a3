rm -f gpurun_out/r2_tune9.log
for spec in "city10k long" "mega200k short" "city10k short"; do set -- $spec; timeout 300 python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 --slots 2,4,6 >> gpurun_out/r2_tune9.log 2>&1; done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune9.log'):
    try: d=json.loads(ln)
    except: continue
    print(d['map'],d['kind'],d['slots'],d['smem'],d['ms'],d['Mtraces_s'])
PY
bash tools/prof.sh stshort --map city10k --kind short --n 16777216

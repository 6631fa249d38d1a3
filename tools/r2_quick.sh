# quick A/B of the working tree: the three GLOBAL workloads + rooms8
rm -f gpurun_out/r2_quick.log
for spec in "city10k long" "mega200k short" "city10k short" "rooms8 uniform" "city10k uniform" "mega200k long"; do set -- $spec
  python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 $EXTRA 2>&1 | tail -1 >> gpurun_out/r2_quick.log
done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_quick.log'):
    try: d=json.loads(ln)
    except: print(ln.strip()[:200]); continue
    print(d['map'],d['kind'],'variant',d['variant'],d['ms'],d['Mtraces_s'],d['hits'])
PY

# box phase / general phase split: parity, then thresholds
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "parity_vs_oracle or knobs or extreme or odd or mega" > gpurun_out/r2_t26.log 2>&1; tail -2 gpurun_out/r2_t26.log
rm -f gpurun_out/r2_tune26.log
for gt in 4 8 12 16; do for bs in 1 2; do
for spec in "city10k long" "rooms8 uniform" "mega200k short"; do set -- $spec
  BSPGPU_GT=$gt BSPGPU_BS=$bs python tools/tune.py --map $1 --kind $2 --n 33554432 --reps 3 2>&1 | tail -1 | sed "s/^/gt$gt bs$bs /" >> gpurun_out/r2_tune26.log
done; done; done
python - <<'PY'
import json
for ln in open('gpurun_out/r2_tune26.log'):
    a,b,js=ln.split(' ',2)
    try: d=json.loads(js)
    except: print(ln[:200]); continue
    print(a,b,d['map'],d['kind'],d['ms'],d['Mtraces_s'])
PY

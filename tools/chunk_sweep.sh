for c in 262144 524288 1048576 2097152 4194304; do
  python bench.py --steps 3 --warmup 3 --no-cpu --segments-per-gpu 33554432 --chunk $c 2>/dev/null > /tmp/b.json
  python -c "import json; d=json.loads(open('/tmp/b.json').read().strip().splitlines()[-1]); print($c, 'e2e %.4e' % d['e2e']['value'])"
done

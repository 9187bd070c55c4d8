#!/bin/bash
# C2 bundled-scene row: parity test + bench leg
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_compositor.py -x -q -m gpu 2>&1 | tail -15
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_c2.json 2> gpurun_out/bench_c2.err
echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open('gpurun_out/bench_c2.json').read().strip().splitlines()[-1])
print(json.dumps(d.get('splat_step', {}).get('c2_bundled'), indent=1))
print(d['value'], d['roofline'])
PY
tail -5 gpurun_out/bench_c2.err

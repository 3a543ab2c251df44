#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?"; tail -12 gpurun_out/${TAG}_tests.log
python tools/hang_stress.py --seconds 25 --jitter-ms 1 2>&1 | tail -2
for P in 1 2 3; do
  timeout 600 python bench.py --steps 100 --warmup 5 --plans $P --no-extras --no-cpu-baseline > gpurun_out/${TAG}_plans$P.json 2> gpurun_out/${TAG}_plans$P.err
  python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_plans$P.json').read().strip().splitlines()[-1])
print('plans $P value', round(d['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call']['value']), 'clk', d['clocks']['sm_mhz'])
print(' hbm', [(r['kernel'], round(r['us'],1), round(r['frac'],3)) for r in d['hbm_roofline']])
PY
done

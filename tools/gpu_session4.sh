#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for CFG in yolov3_416 spp_608; do
timeout 600 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline --config $CFG > gpurun_out/${TAG}_bench_$CFG.json 2> gpurun_out/${TAG}_bench_$CFG.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench_$CFG.json').read().strip().splitlines()[-1])
print('$CFG value', round(d['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call']['value']), 'clk', d['clocks']['sm_mhz'], 'conv_seq_ms', round(d['roofline']['conv_ms_per_step'],3), 'ms/step', round(d['ms_per_step'],3))
PY
done
timeout 300 python bench.py --config nms_stress --steps 20 | cut -c1-700

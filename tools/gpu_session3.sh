#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 900 python -m pytest tests/test_gpu_round2.py -m gpu -q -k "resident or recycled or spp3" > gpurun_out/${TAG}_tests_new.log 2>&1
echo "new tests rc=$?"; tail -6 gpurun_out/${TAG}_tests_new.log
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/${TAG}_tests.log 2>&1
echo "all tests rc=$?"; tail -4 gpurun_out/${TAG}_tests.log
for V in 0 1; do
  Y3_NO_BRES=$V timeout 600 python tools/conv_report.py > gpurun_out/${TAG}_conv_report_nobres$V.txt 2>&1
  head -3 gpurun_out/${TAG}_conv_report_nobres$V.txt
done
timeout 600 python bench.py --steps 100 --warmup 5 --no-extras --no-cpu-baseline > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
python - <<PY
import json
d=json.loads(open('gpurun_out/${TAG}_bench.json').read().strip().splitlines()[-1])
print('value', round(d['value']), 'sustained', round(d['sustained']['value']), 'e2e', round(d['e2e']['value']), 'sync', round(d['e2e']['sync_call']['value']), 'clk', d['clocks']['sm_mhz'], 'conv_seq_ms', round(d['roofline']['conv_ms_per_step'],3), 'alone', round(d['roofline']['kernels_alone']['conv_ms_per_step'],3))
PY

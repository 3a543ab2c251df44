#!/bin/bash
# One GPU-box session: GPU tests, then the default bench line (what the driver runs).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 1500 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/${TAG}_tests.log 2>&1
echo "tests rc=$?" | tee -a gpurun_out/${TAG}_tests.log
tail -5 gpurun_out/${TAG}_tests.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err
echo "bench rc=$?"
tail -c 1500 gpurun_out/${TAG}_bench.err
head -c 3000 gpurun_out/${TAG}_bench.json

#!/bin/bash
# A/B of inference()'s sub-batch pipelining (run on the GPU box): e2e images/s for 1, 2 and 4 sub-batches.
for n in 1 2 4; do
  Y3_SUB_BATCHES=$n python bench.py --steps 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('sub_batches', $n, 'device', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done

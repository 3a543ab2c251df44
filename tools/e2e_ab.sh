#!/bin/bash
# A/B of inference()'s sub-batch pipelining (run on the GPU box): e2e images/s for several span layouts.
for v in "Y3_SUB_BATCHES=1" "Y3_SUB_BATCHES=2" "Y3_SUB_BATCHES=4" "Y3_SUB_SPLIT=1,3,4" "Y3_SUB_SPLIT=1,2,2,3" "Y3_SUB_SPLIT=1,7" "Y3_SUB_SPLIT=2,6,8" "Y3_SUB_SPLIT=1,3,4"; do
  env $v python bench.py --steps 30 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', 'device', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done

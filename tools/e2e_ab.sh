#!/bin/bash
# A/B of inference()'s sub-batch pipelining (run on the GPU box): e2e images/s for several span layouts.
# usage: tools/e2e_ab.sh "Y3_SUB_SPLIT=1,3,4" "Y3_SUB_BATCHES=4" ...
for v in "$@"; do
  env $v python bench.py --steps 20 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('$v', 'device', round(d['value']), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value']))"
done

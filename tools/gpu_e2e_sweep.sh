#!/bin/bash
N=${N:-8}
run() { env "$@" python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600 + RANDOM % 300)) tools/e2e_probe.py --tag "$*" $EXTRA 2>/dev/null | tail -1; }
nproc
run Y3_PROBE_PINNED=1
EXTRA=--no-gather run Y3_PROBE_PINNED=1

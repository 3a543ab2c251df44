#!/bin/bash
# A/B of library builds / knobs on ONE box: device step time of the benchmark workload (graph replay).
# usage: tools/ab_step.sh "ENV=.. ENV2=.." "ENV=.." ...   (each argument = one variant's environment)
for round in 1 2; do
  for v in "$@"; do
    env $v python tools/overlap_probe.py --splits 1 --iters 40 2>/dev/null | sed "s|^|[$v] |"
  done
done

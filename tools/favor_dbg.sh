#!/bin/bash
# timing experiments: which stage of the FAVOR pipeline bounds the pair shape
for d in ${@:-0 7 8 15}; do
  echo "dbg=$d"; RFK_FAVOR_DBG=$d timeout 40 python tools/bench_favor.py 2>&1 | tail -3
done

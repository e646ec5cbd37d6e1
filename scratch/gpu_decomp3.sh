#!/bin/bash
for m in 0x0 0x1800; do
  echo "== mode $m"
  PGM_DEBUG_PROF=1 PGM_DEBUG_MODE=$m python scratch/gpu_decomp.py $m 2>&1 | grep -v "^$" | tail -2
done

#!/bin/bash
for a in 0 1000; do
  echo "== PGM_STAGED_CHOL_ALL_N=$a"
  PGM_STAGED_CHOL_ALL_N=$a timeout 120 python scratch/gpu_large_time.py 2k 16k 2>&1 | tail -4
done
timeout 100 python scratch/gpu_c1_route.py 2>&1 | tail -1

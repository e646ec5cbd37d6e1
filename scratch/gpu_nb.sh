#!/bin/bash
for nb in 48 64; do
  echo "== PGM_STAGED_NB=$nb"
  PGM_STAGED_NB=$nb python scratch/gpu_large_time.py c4 2>&1 | tail -2
done

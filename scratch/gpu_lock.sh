#!/bin/bash
for lk in 0 1; do
  echo "== PGM_PIPE_LOCK=$lk"
  PGM_PIPE_LOCK=$lk python scratch/gpu_time.py 2>&1 | grep "^time"
done

#!/bin/bash
# de-phasing experiment: delay the second co-resident block of every SM by PGM_PHASE_NS
for ns in 0 2000 5000 10000 20000 50000 100000 200000 500000 1000000; do
  echo "== PGM_PHASE_NS=$ns"
  PGM_PHASE_NS=$ns python scratch/gpu_time.py 2>&1 | grep "^time"
done

#!/bin/bash
# hook build, no PGM_DEBUG_PROF: event timings are clean.  skeleton = 0x1f00 (no loads, MMAs, epilogue math, potrf)
for m in 0x1f00 0x3f00 0x5f00 0x9f00 0xff00 0x2000 0x4000; do
  PGM_DEBUG_MODE=$m python scratch/gpu_decomp.py $m 2>&1 | grep "grad=1"
done

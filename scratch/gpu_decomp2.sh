#!/bin/bash
# per-phase decomposition with the PGM_DEBUG_HOOKS build (results are garbage in the switched modes)
for m in 0x0 0x200 0xc00 0x1800 0xe00 0x1e00; do
  echo "== mode $m" 
  PGM_DEBUG_PROF=1 PGM_DEBUG_MODE=$m python scratch/gpu_decomp.py $m 2>&1 | grep -v "^$" | tail -4
done
echo "== one block per SM, mode 0"
PGM_DEBUG_SMEM_KB=120 PGM_DEBUG_MODE=0 python scratch/gpu_decomp.py 0 2>&1 | tail -2

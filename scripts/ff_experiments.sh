#!/bin/bash
# timing experiments on the fused forward kernel: what is left out (B200_FF_DBG bits: 1 no MMA, 2 no flush, 4 no skin, 8 no reloads)
for d in 0 1 2 4 8 6 3 7; do
  echo -n "dbg=$d "; B200_FF_DBG=$d timeout 100 python scripts/time_kernels.py fp32 4096 2>&1 | grep -o "blend_lbs_fwd=[0-9]*us"
done



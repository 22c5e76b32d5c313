#!/bin/bash
# config 5 through the time-batched route: shape of the EQ pre-pass (streams per CTA, shared memory asked for) x blocks per call
for gs in "3 0" "6 0" "6 100" "6 160" "6 200"; do
  set -- $gs
  for k in 64 256; do
    echo "G=$1 smem=$2 K=$k: $(OHS_TB_EQ_G=$1 OHS_TB_EQ_SMEM_KB=$2 python tools/profile_cfg5_batched.py $k 2>&1 | tail -1)"
  done
done

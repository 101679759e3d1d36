#!/bin/bash
# A/B of the segment schedule (developer tool)
for v in 0 1 64; do
  echo "== MQVS_SEG_ONE_PER_CTA=$v"
  for cfg in "1000000 fp32 64 10" "1000000 bf16 64 10" "1000000 fp32 8 10" "1000000 bf16 256 10" "1000000 bf16 4096 10" "12500000 bf16 64 10"; do
    MQVS_SEG_ONE_PER_CTA=$v timeout -s KILL 100 python tools/tc_one.py $cfg 0 6 2>&1 | tail -1 | sed "s/^/$cfg : /" | cut -c1-140
  done
done

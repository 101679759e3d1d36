#!/bin/bash
# A/B of the segment schedule (developer tool)
for g in 16 64 100000; do
  echo "== MQVS_SEG_GROWTH_B128=$g"
  MQVS_SEG_GROWTH_B128=$g timeout -s KILL 100 python tools/tc_one.py 1000000 fp32 64 10 0 6 2>&1 | tail -1
  MQVS_SEG_GROWTH_B128=$g timeout -s KILL 100 python tools/tc_one.py 1000000 bf16 64 10 0 6 2>&1 | tail -1
done
for g in 8 16 32 64; do
  echo "== MQVS_SEG_GROWTH=$g"
  MQVS_SEG_GROWTH=$g timeout -s KILL 100 python tools/tc_one.py 1000000 bf16 4096 10 0 4 2>&1 | tail -1
  MQVS_SEG_GROWTH=$g timeout -s KILL 200 python tools/tc_one.py 10000000 bf16 4096 100 0 3 2>&1 | tail -1
done

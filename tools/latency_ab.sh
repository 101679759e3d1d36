#!/bin/bash
# Latency path A/B on one box: host microseconds per blocking single-query search (tools/latency_one.py), with the spin wait on
# the mapped completion word and the timed CUDA events switched on / off.   bash tools/latency_ab.sh > gpurun_out/latency_ab.log
for n in 154 4096 65536; do
  for dt in fp32 bf16; do
    for rep in 1 2; do
      echo "rows=$n $dt default            : $(python tools/latency_one.py $n $dt 2000 | tail -1)"
      echo "rows=$n $dt no spin wait       : $(MQVS_NO_SPIN_WAIT=1 python tools/latency_one.py $n $dt 2000 | tail -1)"
      echo "rows=$n $dt no events          : $(MQVS_TIMING=0 python tools/latency_one.py $n $dt 2000 | tail -1)"
      echo "rows=$n $dt no events, no spin : $(MQVS_TIMING=0 MQVS_NO_SPIN_WAIT=1 python tools/latency_one.py $n $dt 2000 | tail -1)"
    done
  done
done

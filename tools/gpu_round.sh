#!/bin/bash
# One gpurun call's worth of checks; every step logs to gpurun_out/ and the script goes on whatever fails.
#   tools/gpu_round.sh [tag] [steps...]   steps: tests bench c2cmp ingest  (default: all)
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r02}; shift
STEPS=${*:-tests bench c2cmp ingest}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.txt 2>&1
for s in $STEPS; do
  case $s in
    pairsanity) timeout 200 python tools/tc_one.py 300000 bf16 512 10 2 > $OUT/${TAG}_pairsanity.log 2>&1; rc=$?; echo "pairsanity rc=$rc"
            tail -2 $OUT/${TAG}_pairsanity.log
            if [ $rc -ne 0 ] || ! grep -q "planted ok True" $OUT/${TAG}_pairsanity.log; then export MQVS_NO_PAIR=1; echo "PAIR KERNEL DISABLED for the rest of this call"; fi ;;
    tests)  timeout 1500 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 -p no:cacheprovider > $OUT/${TAG}_pytest_gpu.log 2>&1; echo "tests rc=$?" ;;
    ktests) timeout 900 python -m pytest tests -m gpu -q --maxfail=25 --timeout 600 -p no:cacheprovider -k "$KEXPR" > $OUT/${TAG}_pytest_gpu_k.log 2>&1; echo "ktests rc=$?"; tail -15 $OUT/${TAG}_pytest_gpu_k.log ;;
    bench)  timeout 900 python bench.py --steps 10 --warmup 3 > $OUT/${TAG}_bench_n1.json 2> $OUT/${TAG}_bench_n1.err; echo "bench rc=$?" ;;
    c2cmp)  for v in half full; do
              if [ $v = full ]; then export MQVS_NO_HALF_TILE=1; else unset MQVS_NO_HALF_TILE; fi
              timeout 300 python bench.py --workload c2 --no-secondary --no-cpu-baseline --points 1,16,128 --steps 30 --warmup 5 \
                  > $OUT/${TAG}_bench_c2_${v}.json 2> $OUT/${TAG}_bench_c2_${v}.err; echo "c2 $v rc=$?"
              timeout 300 python bench.py --workload c4 --rows 12500000 --no-secondary --no-cpu-baseline --points 1,128,256,1024 --steps 20 --warmup 5 \
                  > $OUT/${TAG}_bench_c4shard_${v}.json 2> $OUT/${TAG}_bench_c4shard_${v}.err; echo "c4 shard $v rc=$?"
            done; unset MQVS_NO_HALF_TILE ;;
    ingest) timeout 300 python tools/ingest_check.py 1000000 > $OUT/${TAG}_ingest.log 2>&1; echo "ingest rc=$?"; cp $OUT/ingest_check.json $OUT/${TAG}_ingest_check.json 2>/dev/null ;;
    e2e)    timeout 300 python tools/e2e_breakdown.py 1000000 64 fp32 > $OUT/${TAG}_e2e_c2.log 2>&1; echo "e2e rc=$?"
            timeout 300 python tools/e2e_breakdown.py 1000000 1 fp32 > $OUT/${TAG}_e2e_c2_b1.log 2>&1 ;;
    paircmp) for v in pair nopair; do
              if [ $v = nopair ]; then export MQVS_NO_PAIR=1; else unset MQVS_NO_PAIR; fi
              timeout 400 python bench.py --workload c4 --rows 12500000 --batch 1024 --no-secondary --no-cpu-baseline --points 512,4096 --steps 10 --warmup 3 \
                  > $OUT/${TAG}_bench_c4shard_${v}.json 2> $OUT/${TAG}_bench_c4shard_${v}.err; echo "c4 shard $v rc=$?"
              timeout 400 python bench.py --workload c3 --no-secondary --no-cpu-baseline --steps 5 --warmup 3 \
                  > $OUT/${TAG}_bench_c3_${v}.json 2> $OUT/${TAG}_bench_c3_${v}.err; echo "c3 $v rc=$?"
              MQVS_TRACE=1 timeout 300 python tools/tc_one.py 12500000 bf16 256 10 4 > $OUT/${TAG}_trace_b256_seg_${v}.log 2>&1
            done; unset MQVS_NO_PAIR ;;
    ncu)    # launch list of the default bench (headline only) + full captures of the dominant kernels; plain run first
            CMD="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-store-e2e --points 1"
            timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $OUT/${TAG}_ncu_launches_bench_default.csv $CMD > $OUT/${TAG}_ncu_launches.log 2>&1
            echo "ncu launches rc=$?"
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fused_kernel -s 3 -c 1 -o $OUT/${TAG}_prof_c4_b64_fused $CMD > $OUT/${TAG}_ncu_fused.log 2>&1
            echo "ncu fused rc=$?"
            for B in 256 1024; do
              CMD2="python tools/tc_one.py 12500000 bf16 $B 10 0 2"
              timeout 300 $CMD2 > $OUT/${TAG}_ncu_plain_b$B.log 2>&1 &&
              timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_score_pair -s 3 -c 1 -o $OUT/${TAG}_prof_pair_b$B $CMD2 > $OUT/${TAG}_ncu_pair_b$B.log 2>&1
              echo "ncu pair B=$B rc=$?"
            done ;;
    ncufused) CMD="python bench.py --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --no-store-e2e --points 1"
            timeout 300 $CMD > $OUT/${TAG}_ncu_plain.log 2>&1 &&
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fused_kernel -s 2 -c 2 -o $OUT/${TAG}_prof_c4_b64_fused $CMD > $OUT/${TAG}_ncu_fused.log 2>&1
            echo "ncu fused rc=$?" ;;
    shadowx) timeout 600 python tools/shadow_crossover.py > $OUT/${TAG}_shadow_crossover.json 2> $OUT/${TAG}_shadow_crossover.err; echo "shadowx rc=$?" ;;
    ncuc2)  CMD="python bench.py --workload c2 --steps 2 --warmup 3 --no-secondary --no-cpu-baseline --points 1"
            timeout 300 $CMD > $OUT/${TAG}_ncu_c2_plain.log 2>&1 &&
            timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $OUT/${TAG}_ncu_launches_c2.csv $CMD > $OUT/${TAG}_ncu_launches_c2.log 2>&1
            echo "ncu c2 launches rc=$?"
            timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_fused_kernel -s 2 -c 2 -o $OUT/${TAG}_prof_c2_b64_fused $CMD > $OUT/${TAG}_ncu_c2_fused.log 2>&1
            echo "ncu c2 fused rc=$?" ;;
    c2)     timeout 300 python bench.py --workload c2 --no-secondary --no-cpu-baseline --points 1,16,256,1024 --steps 30 --warmup 5 > $OUT/${TAG}_bench_c2.json 2> $OUT/${TAG}_bench_c2.err; echo "c2 rc=$?"
            MQVS_NO_SHADOW=1 timeout 300 python bench.py --workload c2 --no-secondary --no-cpu-baseline --points 1,16,256,1024 --steps 30 --warmup 5 > $OUT/${TAG}_bench_c2_noshadow.json 2> $OUT/${TAG}_bench_c2_noshadow.err; echo "c2 noshadow rc=$?" ;;
    tracec2) MQVS_TRACE=1 timeout 200 python tools/tc_one.py 1000000 fp32 64 10 0 4 > $OUT/${TAG}_trace_c2_b64.log 2>&1; echo "trace rc=$?"
            MQVS_TRACE=1 timeout 200 python tools/tc_one.py 1000000 fp32 1 10 0 4 > $OUT/${TAG}_trace_c2_b1.log 2>&1 ;;
    latency) for r in 154 10000 100000; do
              timeout 120 python tools/latency_one.py $r fp32 3000 > $OUT/${TAG}_latency_${r}.log 2>&1; tail -1 $OUT/${TAG}_latency_${r}.log
              MQVS_NO_INLINE_QUERIES=1 timeout 120 python tools/latency_one.py $r fp32 3000 > $OUT/${TAG}_latency_${r}_copied.log 2>&1; tail -1 $OUT/${TAG}_latency_${r}_copied.log
            done ;;
    c3ip)   timeout 400 python bench.py --workload c3 --space ip --no-secondary --no-cpu-baseline --steps 5 --warmup 3 > $OUT/${TAG}_bench_c3_ip.json 2> $OUT/${TAG}_bench_c3_ip.err; echo "c3 ip rc=$?" ;;
    widek)  for kk in 100 200 1000; do timeout 200 python tools/tc_one.py 1000000 fp32 64 $kk 0 5 > $OUT/${TAG}_widek_$kk.log 2>&1; tail -1 $OUT/${TAG}_widek_$kk.log; done ;;
    smoke)  timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?" ;;
  esac
done
tail -5 $OUT/${TAG}_pytest_gpu.log 2>/dev/null
exit 0

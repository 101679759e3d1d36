#!/bin/bash
# A/B of two builds of libmqvs.so on ONE box: tools/ab_lib.sh <tag> <before.so> -- runs the same bench commands with the
# library at HEAD ("after") and with <before.so> swapped in ("before"), alternating so that both see the same clock state.
cd "$(dirname "$0")/.." || exit 1
TAG=$1; BEFORE=$2; export ABTAG=$TAG; OUT=gpurun_out; mkdir -p $OUT
LIB=mediquery-rag_b200/libmqvs.so
cp $LIB /tmp/libmqvs_after.so
run() {  # run <name> <bench args...>
    name=$1; shift
    for rep in 1 2; do
        for which in after before; do
            cp /tmp/libmqvs_$which.so $LIB 2>/dev/null || cp $BEFORE $LIB
            [ $which = before ] && cp $BEFORE $LIB
            timeout 600 python bench.py "$@" --no-cpu-baseline --no-secondary > $OUT/${TAG}_${name}_${which}_$rep.json 2> $OUT/${TAG}_${name}_${which}_$rep.err
            echo "$name $which $rep rc=$?"
        done
    done
}
if [ -n "$AB_RUNS" ]; then  # AB_RUNS="name1:bench args;name2:bench args"
    IFS=';' read -ra SPECS <<< "$AB_RUNS"
    for spec in "${SPECS[@]}"; do run "${spec%%:*}" ${spec#*:}; done
else
    run c4n1 --steps 10 --warmup 3 --points 1
    run c4shard --rows 12500000 --steps 20 --warmup 5 --points 1,16
    run c2 --workload c2 --steps 20 --warmup 5 --points 1
fi
cp /tmp/libmqvs_after.so $LIB
python - <<'P'
import glob, json
for f in sorted(glob.glob("gpurun_out/%s_*.json" % __import__("os").environ.get("ABTAG", ""))):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except Exception as e:
        print(f, "unreadable", e); continue
    pts = " ".join("B=%d %.4f ms (%s MHz)" % (p["batch"], p["ms_per_step"], p["clocks"]["sm_mhz"]) for p in d.get("points", []))
    print("%-52s B=%d %.4f ms kernel %.4f ms (%s MHz) | %s" % (f.split("/")[-1], d["config"]["batch"], d["ms_per_step"], d["roofline"]["avg_launch_ms"], d["clocks"]["sm_mhz"], pts))
P

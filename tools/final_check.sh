#!/bin/bash
# One box, HEAD: latency-path A/B, pytest -m gpu, smoke(), both bench arms.   tools/final_check.sh [tag]   (logs -> gpurun_out/<tag>_*)
cd "$(dirname "$0")/.." || exit 1
T=${1:-final}; OUT=gpurun_out; mkdir -p $OUT
for v in "" "MQVS_NO_SPIN_WAIT=1" "MQVS_TIMING=2" "MQVS_TIMING=0"; do
  for n in 154 4096; do echo "rows=$n fp32 [$v]: $(env $v python tools/latency_one.py $n fp32 2000 | tail -1)"; done
done > $OUT/${T}_latency_ab.log 2>&1
cat $OUT/${T}_latency_ab.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > $OUT/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 $OUT/${T}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference > $OUT/${T}_bench_ref.json 2> $OUT/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > $OUT/${T}_bench_default.json 2> $OUT/${T}_bench_default.err; echo "bench rc=$?"
python - "$OUT/${T}_bench_default.json" <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "e2e", d["e2e"]["value"], "roofline.frac", d["roofline"]["frac"], "clocks", d["clocks"],
      "config1 us/search", d["secondary"]["config1"].get("e2e_us_per_search"))
PY

T=r02y
for v in "" "MQVS_NO_SPIN_WAIT=1" "MQVS_TIMING=2" "MQVS_TIMING=0"; do for n in 154 4096; do echo "rows=$n fp32 [$v]: $(env $v python tools/latency_one.py $n fp32 2000 | tail -1)"; done; done > gpurun_out/${T}_latency_ab.log 2>&1
cat gpurun_out/${T}_latency_ab.log
timeout 900 python -m pytest tests -m gpu -x -q -p no:cacheprovider > gpurun_out/${T}_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/${T}_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --impl reference > gpurun_out/${T}_bench_ref.json 2> gpurun_out/${T}_bench_ref.err; echo "ref rc=$?"
timeout 600 python bench.py > gpurun_out/${T}_bench_default.json 2> gpurun_out/${T}_bench_default.err; echo "bench rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r02y_bench_default.json').read().strip().splitlines()[-1])
print(d['value'], d['e2e']['value'], d['roofline']['frac'], d['clocks'], d['secondary']['config1'].get('e2e_us_per_search'))
PY

#!/bin/bash
# multi-GPU round: tools/gpu_round2.sh <tag> <N>
cd "$(dirname "$0")/.." || exit 1
TAG=${1:-r02}; N=${2:-2}
OUT=gpurun_out; mkdir -p $OUT
nvidia-smi --query-gpu=index,name,clocks.sm,power.draw,memory.used --format=csv > $OUT/${TAG}_smi.txt 2>&1
nvidia-smi topo -m > $OUT/${TAG}_topo.txt 2>&1
timeout 600 python tools/multi_device_parity.py $N > $OUT/${TAG}_multi_device_parity.log 2>&1; echo "multi_device_parity rc=$?"; tail -3 $OUT/${TAG}_multi_device_parity.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tools/multi_gpu_parity.py > $OUT/${TAG}_multi_gpu_parity.log 2>&1; echo "multi_gpu_parity rc=$?"; tail -3 $OUT/${TAG}_multi_gpu_parity.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.json 2> $OUT/${TAG}_bench_n$N.err; echo "bench n$N rc=$?"; tail -c 600 $OUT/${TAG}_bench_n$N.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29535 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $OUT/${TAG}_bench_ref_n$N.json 2> $OUT/${TAG}_bench_ref_n$N.err; echo "ref n$N rc=$?"

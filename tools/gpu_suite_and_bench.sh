#!/bin/bash
# gpurun --gpus 2 script: full GPU suite (multi-GPU tests included), bench at N=1 and N=2 (torchrun).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/final2_timeline.log; }
nvidia-smi -L > $O/final2_gpus.log 2>&1
timeout 300 python -m pytest tests -x -q -m gpu > $O/final2_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/final2_pytest_gpu.log)"
timeout 200 python bench.py > $O/final2_bench_n1.json 2> $O/final2_bench_n1.err; el "bench n1 rc=$?"
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 5 --warmup 3 > $O/final2_bench_n2.json 2> $O/final2_bench_n2.err; el "bench n2 rc=$?"
el done

#!/bin/bash
# gpurun --gpus 2 script: the round-2 kernels at two ranks (torchrun, NCCL): multi-device test, C2 float with the shard check, configs[2].
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s18_timeline.log; }
timeout 300 python -m pytest tests/test_gpu_multi.py -q -m gpu -x > $O/s18_pytest_multi.log 2>&1; el "pytest multi rc=$?: $(tail -1 $O/s18_pytest_multi.log)"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 600 $TR bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu --check-shards > $O/s18_bench_c2_n2.json 2> $O/s18_bench_c2_n2.err; el "bench c2 n=2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c3_svy_spi_wt_2e6_2e7 --steps 2 --warmup 1 --no-cpu > $O/s18_bench_c3_n2.json 2> $O/s18_bench_c3_n2.err; el "bench c3 n=2 rc=$?"
timeout 600 $TR bench.py --gpus 2 --workload c4_box_smu_clustered_1e7 --steps 3 --warmup 3 --no-cpu --check-shards > $O/s18_bench_c4_n2.json 2> $O/s18_bench_c4_n2.err; el "bench c4 n=2 rc=$?"
python - <<'PY' | tee -a gpurun_out/s18_timeline.log
import json
for t in ("c2", "c3", "c4"):
    try:
        d = json.loads(open(f"gpurun_out/s18_bench_{t}_n2.json").read().strip().splitlines()[-1])
        print(t, "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "balance", d.get("load_balance_max_over_mean"), "shards", d.get("shard_parity"), "parity", d.get("parity_check"))
    except Exception as ex:
        print(t, "no line:", ex)
PY
el done

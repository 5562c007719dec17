#!/bin/bash
# gpurun --gpus 8 script: the round-2 kernels on one 8xB200 box: configs[1] at 8 and 4 ranks, configs[3] and configs[4] at size.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s20_timeline.log; }
run() {  # ranks tag args...
  local n=$1 tag=$2; shift 2
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600 + n)) bench.py --gpus $n "$@" > $O/s20_bench_${tag}_n$n.json 2> $O/s20_bench_${tag}_n$n.err
  el "bench $tag n=$n rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("$O/s20_bench_${tag}_n$n.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "balance", round(d.get("load_balance_max_over_mean", 0), 4), "shards", d.get("shard_parity"), "hbm", d.get("hbm_used_gb_per_rank", [None])[0])
except Exception as ex:
    print("no line:", ex)
PY
)"
}
run 8 c2_float --steps 5 --warmup 3 --no-cpu --check-shards
run 4 c2_float --steps 5 --warmup 3 --no-cpu
run 8 c4_1e8 --workload c4_box_smu_clustered_1e8 --steps 2 --warmup 1 --no-cpu --check-shards
run 8 c5_1e8 --workload c5_svy_spi_wt_2e6_1e8 --steps 2 --warmup 1 --no-cpu --check-shards
el done

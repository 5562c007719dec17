#!/bin/bash
# gpurun --gpus N script: multi-GPU evidence.  usage: gpu_r2_multi.sh <N> [big]
#   bench at 1..N ranks (torchrun, NCCL) on the headline workload, float and double; GPU suite incl. the multi-device test;
#   the real drop-in (unmodified FCFC host + shim) on the 10^7 box at 1 and N devices;
#   with "big": configs[3] and configs[4] at size (10^8 points) on all N GPUs with the shard-decomposition check.
set -u
cd "$(dirname "$0")/../.."
N=${1:-2}; BIG=${2:-}
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/m${N}_timeline.log; }
nvidia-smi -L > $O/m${N}_gpus.log 2>&1; nproc >> $O/m${N}_gpus.log; free -g >> $O/m${N}_gpus.log
run_bench() {  # ranks tag args...
  local n=$1 tag=$2; shift 2
  if [ "$n" = 1 ]; then timeout 600 python bench.py --gpus 1 "$@" > $O/m${N}_bench_${tag}_n1.json 2> $O/m${N}_bench_${tag}_n1.err
  else timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@" > $O/m${N}_bench_${tag}_n$n.json 2> $O/m${N}_bench_${tag}_n$n.err
  fi
  el "bench $tag n=$n rc=$? $(python - <<PY
import json
try:
    d = json.loads(open("$O/m${N}_bench_${tag}_n$n.json").read().strip().splitlines()[-1])
    print("ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "balance", round(d.get("load_balance_max_over_mean", 0), 4), "shards", d.get("shard_parity"), "parity", (d.get("parity_check") or {}).get("exact", (d.get("parity_check") or {}).get("within_reference_spread")))
except Exception as ex:
    print("no line:", ex)
PY
)"
}
timeout 600 python -m pytest tests/test_gpu_multi.py tests/test_gpu_cnvt.py tests/test_gpu_parity.py -q -m gpu -x -s > $O/m${N}_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/m${N}_pytest.log)"
for n in 1 2 4 8; do
  [ $n -le $N ] || continue
  [ -z "${SKIP_BENCH:-}" ] || continue
  run_bench $n c2_float --steps 5 --warmup 3 --no-cpu --check-shards
  run_bench $n c2_double --steps 5 --warmup 3 --no-cpu --prec double
done
# the real host: FCFC_2PT_BOX (unmodified sources + shim) on the 10^7 box, 1 device and all devices
if [ -x integration/_build/flt/FCFC_2PT_BOX ]; then
  python tools/make_box_ascii.py 10000000 2000 $O/box1e7.txt > /dev/null 2>&1; el "ascii catalogue written"
  for dev in 0 all; do
    d=$O/m${N}_cli_$dev; mkdir -p $d
    cat > $d/fcfc.conf <<CONF
CATALOG = "$O/box1e7.txt"
CATALOG_LABEL = D
ASCII_FORMATTER = "%f %f %f"
POSITION = ["\$1","\$2","\$3"]
BOX_SIZE = 2000
BINNING_SCHEME = 1
PAIR_COUNT = DD
PAIR_COUNT_FILE = "$d/DD.bin"
CF_ESTIMATOR = "DD / @@ - 1"
CF_OUTPUT_FILE = "$d/xi.txt"
MULTIPOLE = [0,2,4]
MULTIPOLE_FILE = "$d/xil.txt"
SEP_BIN_MIN = 0
SEP_BIN_MAX = 200
SEP_BIN_SIZE = 5
MU_BIN_NUM = 120
OUTPUT_FORMAT = 0
OVERWRITE = 2
VERBOSE = T
CONF
    if [ $dev = 0 ]; then export FCFC_GPU_DEVICES=0; else unset FCFC_GPU_DEVICES; fi
    ( time FCFC_GPU_ARITH=fma FCFC_GPU_VERBOSE=1 OMP_NUM_THREADS=$(nproc) timeout 900 integration/_build/flt/FCFC_2PT_BOX -c $d/fcfc.conf ) > $d/run.log 2>&1
    el "FCFC_2PT_BOX devices=$dev rc=$? $(grep -E 'count step|real' $d/run.log | tr '\n' ' ')"
  done
  unset FCFC_GPU_DEVICES
  cmp $O/m${N}_cli_0/DD.bin $O/m${N}_cli_all/DD.bin && el "DD.bin identical at 1 and $N devices"
  rm -f $O/box1e7.txt $O/m${N}_cli_*/DD.bin
fi
if [ -n "$BIG" ]; then
  run_bench $N c4_1e8 --workload c4_box_smu_clustered_1e8 --steps 2 --warmup 1 --no-cpu --check-shards
  run_bench $N c5_1e8 --workload c5_svy_spi_wt_2e6_1e8 --steps 2 --warmup 1 --no-cpu --check-shards
fi
el done

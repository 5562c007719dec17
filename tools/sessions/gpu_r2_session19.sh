#!/bin/bash
# gpurun script: tiles of 96 points for count_kernel_cl: full GPU suite, A/B of the options, default bench line.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s19_timeline.log; }
timeout 900 python -m pytest tests -q -m gpu -x > $O/s19_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s19_pytest_gpu.log)"
timeout 600 python tools/time_opts.py "no_classify=1" "no_classify=0" > $O/s19_time_opts.log 2>&1; el "time_opts rc=$?"; cat $O/s19_time_opts.log | tee -a $O/s19_timeline.log
timeout 300 python bench.py > $O/s19_bench_n1.json 2> $O/s19_bench_n1.err; el "bench rc=$?"
timeout 300 python bench.py --workload c4_box_smu_clustered_1e7 --no-cpu > $O/s19_bench_c4.json 2> $O/s19_bench_c4.err; el "bench c4 rc=$?"
timeout 300 python bench.py --workload c1_box_iso_1e6 --no-cpu > $O/s19_bench_c1.json 2> $O/s19_bench_c1.err; el "bench c1 rc=$?"
python - <<'PY' | tee -a gpurun_out/s19_timeline.log
import json
for t in ("n1", "c4", "c1"):
    try:
        d = json.loads(open(f"gpurun_out/s19_bench_{t}.json").read().strip().splitlines()[-1])
        print(t, "ms/step", round(d["ms_per_step"], 2), "e2e", round(d["e2e"]["ms_per_step"], 2), "frac", d["roofline"]["frac"], "frac_dist", d["roofline"].get("frac_on_distance_evals"), "parity", d.get("parity_check"), "double", d.get("double_precision"))
    except Exception as ex:
        print(t, "no line:", ex)
PY
el done

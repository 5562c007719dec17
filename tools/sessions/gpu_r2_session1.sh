#!/bin/bash
# gpurun script, round 2 / session 1: full GPU suite (with the full-size golden tests), bench lines (float, double,
# reference arm on the real job), A/B of the volatile shared loads, compute-sanitizer logs, ncu launch list.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s1_timeline.log; }
nvidia-smi -L > $O/s1_gpus.log 2>&1; nproc >> $O/s1_gpus.log
timeout 1200 python -m pytest tests -q -m gpu -s > $O/s1_pytest_gpu.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s1_pytest_gpu.log)"
timeout 300 python bench.py > $O/s1_bench_n1.json 2> $O/s1_bench_n1.err; el "bench n1 rc=$?"
timeout 300 python bench.py --prec double --no-cpu > $O/s1_bench_n1_double.json 2> $O/s1_bench_n1_double.err; el "bench n1 double rc=$?"
timeout 300 python bench.py --arith 0 --no-cpu > $O/s1_bench_n1_scalar.json 2> $O/s1_bench_n1_scalar.err; el "bench n1 scalar order rc=$?"
for v in nonvol; do
  if [ -f fcfc_b200/_variants/$v/libfcfc_b200.so ]; then
    for rep in 1 2; do
      timeout 120 python tools/time_c2.py fcfc_b200/libfcfc_b200.so > $O/s1_ab_default.$rep.log 2>&1; el "default: $(grep 'bt=' $O/s1_ab_default.$rep.log | tr '\n' ' ')"
      timeout 120 python tools/time_c2.py fcfc_b200/_variants/$v/libfcfc_b200.so > $O/s1_ab_$v.$rep.log 2>&1; el "$v: $(grep 'bt=' $O/s1_ab_$v.$rep.log | tr '\n' ' ')"
    done
  fi
done
timeout 400 python tools/time_pf.py > $O/s1_time_pf.log 2>&1; el "time_pf rc=$?"; grep same $O/s1_time_pf.log | tee -a $O/s1_timeline.log
timeout 200 python tools/time_tail.py > $O/s1_tail.log 2>&1; el "tail rc=$? $(tail -4 $O/s1_tail.log | tr '\n' ' ')"
timeout 400 python bench.py --impl reference --steps 5 --warmup 0 > $O/s1_bench_reference.json 2> $O/s1_bench_reference.err; el "reference arm rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_case.py 3000 > $O/s1_sanitizer_memcheck.log 2>&1; el "memcheck rc=$?: $(tail -2 $O/s1_sanitizer_memcheck.log | tr '\n' ' ')"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py 1500 > $O/s1_sanitizer_racecheck.log 2>&1; el "racecheck rc=$?: $(tail -2 $O/s1_sanitizer_racecheck.log | tr '\n' ' ')"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/s1_ncu_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > $O/s1_ncu_bench_under_ncu.log 2>&1; el "launch list rc=$?"
el done

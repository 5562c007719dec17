#!/bin/bash
# gpurun script: racecheck after the __syncwarp() around the drains (rotating stack columns); timing check.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s23_timeline.log; }
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_case.py 1500 > $O/s23_sanitizer_racecheck.log 2>&1; el "racecheck rc=$?: $(tail -2 $O/s23_sanitizer_racecheck.log | tr '\n' ' ')"
timeout 600 python tools/time_opts.py "no_classify=0" > $O/s23_time_opts.log 2>&1; el "time_opts rc=$?"; cat $O/s23_time_opts.log | tee -a $O/s23_timeline.log
el done

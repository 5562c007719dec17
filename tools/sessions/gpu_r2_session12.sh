#!/bin/bash
# gpurun script: A/B at 10^7 points of the dense-block variants (old chain / dense_block / 20 warps), with a correctness digest.
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s12_timeline.log; }
timeout 600 python tools/time_c2.py fcfc_b200/libfcfc_b200.so fcfc_b200/_variants/ldrain/libfcfc_b200.so > $O/s12_time_c2.log 2>&1; el "time_c2 rc=$?"
cat $O/s12_time_c2.log | tee -a $O/s12_timeline.log
el done

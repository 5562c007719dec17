#!/bin/bash
# gpurun script: warps per block of the double-precision float-speed kernel (20 / 24 / 28) on the bench workload.
set -u
cd "$(dirname "$0")/.."
O=gpurun_out; mkdir -p $O
timeout 600 python tools/time_c2_double.py fcfc_b200/libfcfc_b200.so fcfc_b200/_variants/df20/libfcfc_b200.so fcfc_b200/_variants/df28/libfcfc_b200.so 2>&1 | tee $O/s24_df_warps.log

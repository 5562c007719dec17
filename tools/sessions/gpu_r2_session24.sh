#!/bin/bash
# gpurun script: cell-size scan at the density of configs[3] at size.
cd "$(dirname "$0")/../.."
timeout 300 python tools/time_density.py 0 4 5 6 7 8 2>&1 | tee gpurun_out/s24_density.log

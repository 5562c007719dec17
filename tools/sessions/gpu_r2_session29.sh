#!/bin/bash
# gpurun --gpus 2 script: the real host (unmodified FCFC + shim) on all devices of the box: CLI drop-in tests, multi-device test,
# and FCFC_2PT_BOX on a 10^7-point ASCII catalogue at 1 and 2 devices (exit status, count step, identical DD.bin).
set -u
cd "$(dirname "$0")/../.."
O=gpurun_out; mkdir -p $O
T0=$(date +%s); el() { echo "[$(( $(date +%s) - T0 )) s] $*" | tee -a $O/s29_timeline.log; }
timeout 600 python -m pytest tests/test_gpu_cli_dropin.py tests/test_gpu_multi.py -q -m gpu > $O/s29_pytest.log 2>&1; el "pytest rc=$?: $(tail -1 $O/s29_pytest.log)"
python tools/make_box_ascii.py 10000000 2000 $O/box1e7.txt > /dev/null 2>&1; el "ascii catalogue written"
for dev in 0 all; do
  d=$O/s29_cli_$dev; mkdir -p $d
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
  el "FCFC_2PT_BOX devices=$dev rc=$? $(grep -E 'count step|resident|real' $d/run.log | tr '\n' ' ')"
done
unset FCFC_GPU_DEVICES
cmp $O/s29_cli_0/DD.bin $O/s29_cli_all/DD.bin && el "DD.bin identical at 1 and 2 devices"
cmp $O/s29_cli_0/xil.txt $O/s29_cli_all/xil.txt && el "multipoles identical at 1 and 2 devices"
rm -f $O/box1e7.txt $O/s29_cli_*/DD.bin
el done

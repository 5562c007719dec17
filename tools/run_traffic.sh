for cb in 1 23; do
FCFC_GPU_COST_BITS=$cb ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,lts__t_sectors_op_read.sum --clock-control none -k regex:count_kernel -s 1 -c 1 --csv --log-file gpurun_out/traffic_cb$cb.csv python tools/prof_one.py 1e7 2000 1 float 1 2 > /dev/null 2>&1
FCFC_GPU_COST_BITS=$cb python tools/time_c2.py fcfc_b200/libfcfc_b200.so 2>&1 | tail -2
done
python tools/time_clustered.py 2>&1 | tail -4

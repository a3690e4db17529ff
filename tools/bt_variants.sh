for l in lib_run1 libhalo_sm100 lib_run4; do echo $l; HALO_B200_LIB=halo_b200/$l.so ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:head_bwd_tc -s 2 -c 1 python tools/bwd_time.py 8 2>&1 | grep -E "duration|dram__" ; done
timeout 300 python -m pytest tests/test_head_gpu.py -m gpu -x -q 2>&1 | tail -2

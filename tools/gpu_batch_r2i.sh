set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2i_tests.log
timeout 600 python bench.py > gpurun_out/r2i_bench_default.json 2> gpurun_out/r2i_bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_|score_|select_|round_|upsample" -c 60 --csv --log-file gpurun_out/launches_r2i_acquire148.csv python tools/profile_step.py --batch 148 --steps 3 --mode acquire > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"head_" -c 60 --csv --log-file gpurun_out/launches_r2i_train8.csv python tools/profile_step.py --batch 8 --steps 3 --mode train > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:head_fwd_tc_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2i_k1 python tools/profile_step.py --batch 32 --steps 2 --mode acquire > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:head_bwd_stream_kernel -s 1 -c 1 -f -o gpurun_out/prof_r2i_k4s python tools/profile_step.py --batch 8 --steps 2 --mode train > /dev/null 2>&1
ls -la gpurun_out/ | tail -8
cat gpurun_out/r2i_tests.log; cat gpurun_out/r2i_bench_default.json | cut -c1-1500

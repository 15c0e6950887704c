#!/bin/bash
# DEV (round 2, GPU call h): GICP with the cluster optimiser, knn5_points, smoke, ncu of the GICP kernels
mkdir -p gpurun_out
( time timeout 200 python -m pytest tests/test_zz_gpu_gicp.py -x -q -m gpu ) > gpurun_out/r2h_gicp.log 2>&1
tail -12 gpurun_out/r2h_gicp.log
( time timeout 120 python -m pytest tests/test_gpu_lio_more.py tests/test_host_cpp.py -q -m gpu -k "neighbour_coordinates or host_adaptors" ) > gpurun_out/r2h_misc.log 2>&1
tail -8 gpurun_out/r2h_misc.log
timeout 120 python tools/gicp_quick.py > gpurun_out/r2h_gicp_quick.json 2> gpurun_out/r2h_gicp_quick.err; cat gpurun_out/r2h_gicp_quick.json; tail -3 gpurun_out/r2h_gicp_quick.err
( time timeout 150 python __graft_entry__.py smoke ) > gpurun_out/r2h_smoke.log 2>&1; tail -4 gpurun_out/r2h_smoke.log
timeout 100 python tools/gicp_prof.py > gpurun_out/r2h_prof_plain.log 2>&1 && \
timeout 200 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 200 --csv --log-file gpurun_out/r2h_gicp_launches.csv python tools/gicp_prof.py > gpurun_out/r2h_ncu_list.log 2>&1 && \
timeout 250 ncu --set full --clock-control none --import-source on -k regex:"k_g_knn_cov|k_g_correspond|k_g_bfgs" -c 4 -o gpurun_out/prof_gicp_r2 -f python tools/gicp_prof.py > gpurun_out/r2h_ncu_full.log 2>&1
cat gpurun_out/r2h_prof_plain.log; tail -2 gpurun_out/r2h_ncu_list.log; tail -2 gpurun_out/r2h_ncu_full.log

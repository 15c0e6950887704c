set -x
python tools/prof_more.py > gpurun_out/prof_more_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_accumulate|k_centroids|k_undistort|k_loam_search|k_loam_accum" -s 6 -c 14 -o gpurun_out/prof_more_r1 -f python tools/prof_more.py > gpurun_out/ncu_full_more.log 2>&1
cat gpurun_out/prof_more_plain.log; tail -3 gpurun_out/ncu_full_more.log

set -x
python tools/prof_all.py > gpurun_out/prof_plain.log 2>&1 && \
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 600 --csv --log-file gpurun_out/r01b_launches.csv python tools/prof_all.py > gpurun_out/ncu_launch.log 2>&1
cat gpurun_out/prof_plain.log
ncu --set full --clock-control none --import-source on -k regex:"k_ndt_eval|k_ndt_score_batch|k_ndt_accumulate|k_ndt_finalize" -s 1 -c 8 -o gpurun_out/prof_ndt_r1 -f python tools/prof_all.py > gpurun_out/ncu_full_ndt.log 2>&1
tail -5 gpurun_out/ncu_full_ndt.log
ncu --set full --clock-control none --import-source on -k regex:"k_search|k_obs" -s 8 -c 4 -o gpurun_out/prof_iekf_r1c -f python tools/prof_all.py > gpurun_out/ncu_full_iekf.log 2>&1
tail -5 gpurun_out/ncu_full_iekf.log
ls -la gpurun_out/*.ncu-rep

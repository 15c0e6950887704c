set -x
export PROF_N_MAP=200000 PROF_N_PRIOR=4000000 PROF_N_HYP=512
python tools/prof_all.py > gpurun_out/prof_plain4.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ndt_score_batch" -s 1 -c 1 -o gpurun_out/prof_score_r1b -f python tools/prof_all.py > gpurun_out/ncu_full_score2.log 2>&1
tail -3 gpurun_out/ncu_full_score2.log

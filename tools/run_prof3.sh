set -x
export SEQ_SCANS=300
python tools/prof_seq.py > gpurun_out/prof_seq_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_search|k_knn5" -s 600 -c 2 -o gpurun_out/prof_search_seq_r1 -f python tools/prof_seq.py > gpurun_out/ncu_full_seq.log 2>&1
cat gpurun_out/prof_seq_plain.log
tail -3 gpurun_out/ncu_full_seq.log

set -x
export PROF_N_MAP=400000 PROF_N_PRIOR=4000000 PROF_N_HYP=256
python tools/prof_all.py > gpurun_out/prof_plain2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_ndt_score_batch" -s 1 -c 1 -o gpurun_out/prof_score_r1 -f python tools/prof_all.py > gpurun_out/ncu_full_score.log 2>&1
tail -3 gpurun_out/ncu_full_score.log
python bench.py --no-ndt --seq-scans 1000 --steps 20 > gpurun_out/bench_seq1000.json 2> gpurun_out/bench_seq1000.err
tail -c 400 gpurun_out/bench_seq1000.err
python -c "
import json; d=json.load(open('gpurun_out/bench_seq1000.json'))
print(json.dumps(d['sequence'])); print(json.dumps(d['roofline']))"

#!/bin/bash
# Round 2, final GPU call: the whole -m gpu suite, the default bench line, then (time permitting) ncu of the TMA-staged search
mkdir -p gpurun_out
( time timeout 420 python -m pytest tests -x -q -m gpu ) > gpurun_out/r2i_pytest.log 2>&1
tail -6 gpurun_out/r2i_pytest.log
( time timeout 300 python bench.py > gpurun_out/r2i_bench.json ) 2> gpurun_out/r2i_bench.err
tail -4 gpurun_out/r2i_bench.err; head -c 600 gpurun_out/r2i_bench.json; echo
B200_KNN_MODE=9 timeout 150 ncu --set full --clock-control none --import-source on -k regex:k_knn5 -s 2 -c 2 -o gpurun_out/prof_knn_tma_r2 -f python tools/knn_prof.py livox > gpurun_out/r2i_ncu_knn_tma.log 2>&1
tail -2 gpurun_out/r2i_ncu_knn_tma.log

#!/bin/bash
# DEV (round 2, GPU call g): GICP parity + TMA-staged k-NN variant + the 32-byte voxel record, short on purpose
mkdir -p gpurun_out
( time timeout 240 python -m pytest tests/test_zz_gpu_gicp.py -x -q -m gpu ) > gpurun_out/r2g_gicp.log 2>&1
tail -15 gpurun_out/r2g_gicp.log
( time timeout 120 python -m pytest tests/test_gpu_voxel.py tests/test_host_cpp.py -q -m gpu -k "builder_matches or capacity or host_adaptors" ) > gpurun_out/r2g_misc.log 2>&1
tail -8 gpurun_out/r2g_misc.log
( B200_KNN_MODE=9 timeout 120 python tests/helpers/knn_mode_worker.py | tail -1 > gpurun_out/r2g_mode9.json; timeout 120 python tests/helpers/knn_mode_worker.py | tail -1 > gpurun_out/r2g_mode7.json ) 2> gpurun_out/r2g_modes.err
cat gpurun_out/r2g_mode9.json gpurun_out/r2g_mode7.json
timeout 120 python tools/gicp_quick.py > gpurun_out/r2g_gicp_quick.json 2> gpurun_out/r2g_gicp_quick.err; cat gpurun_out/r2g_gicp_quick.json; tail -3 gpurun_out/r2g_gicp_quick.err
timeout 150 python tools/knn_quick.py 9 7 > gpurun_out/r2g_knn_quick.log 2>&1; cat gpurun_out/r2g_knn_quick.log

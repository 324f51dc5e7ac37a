#!/bin/bash
# round-1 session-2 call 1: microbenchmarks, C5 breakdown, ncu captures of the current build
set -x
mkdir -p gpurun_out
tools/ubench/dmma_peak > gpurun_out/dmma_peak.log 2>&1
tools/ubench/fp64_lat > gpurun_out/fp64_lat.log 2>&1
python bench.py --steps 100 --warmup 5 --extras c5 --no-cpu > gpurun_out/b_c5.json 2> gpurun_out/b_c5.err
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_driver.py c3 2 > gpurun_out/plain_c3.log 2>&1 && \
  $NCU -k regex:k_gemm_dmma -s 251 -c 1 -f -o gpurun_out/r01_syrk_c3 python tools/prof_driver.py c3 2 > gpurun_out/ncu_c3a.log 2>&1
$NCU -k regex:k_trsm_ll -s 3 -c 1 -f -o gpurun_out/r01_trsm_c3 python tools/prof_driver.py c3 2 > gpurun_out/ncu_c3b.log 2>&1
python tools/prof_driver.py c4 2 > gpurun_out/plain_c4.log 2>&1 && \
  $NCU -k regex:k_ransac_support -s 1 -c 1 -f -o gpurun_out/r01_support_c4 python tools/prof_driver.py c4 2 > gpurun_out/ncu_c4.log 2>&1
python tools/prof_driver.py cublas 3 > gpurun_out/plain_cublas.log 2>&1 && \
  $NCU -k regex:gemm -s 2 -c 1 -f -o gpurun_out/r01_cublas_dgemm python tools/prof_driver.py cublas 3 > gpurun_out/ncu_cublas.log 2>&1
python tools/prof_driver.py c2 6 > gpurun_out/plain_c2.log 2>&1 && \
  ncu --metrics gpu__time_duration.sum --clock-control none -s 44 -c 400 --csv --log-file gpurun_out/r01_c2_launches.csv python tools/prof_driver.py c2 6 > gpurun_out/ncu_c2.log 2>&1
ls -la gpurun_out

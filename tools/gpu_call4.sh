#!/bin/bash
set -x
mkdir -p gpurun_out
RSLAM_LIB=$PWD/ransac_slam_b200/lib/librslam_v8.so python bench.py --steps 20 --warmup 5 --extras c5 --no-cpu > gpurun_out/b_c5_v8.json 2> gpurun_out/b_c5_v8.err
python tools/show_bench.py gpurun_out/b_c5_v8.json 2>/dev/null | grep -A8 "^C5"
NCU="ncu --set full --clock-control none --import-source on"
python tools/prof_driver.py c5 2 > gpurun_out/plain_c5.log 2>&1 && \
  $NCU -k regex:k_syrk_rows -s 3 -c 1 -f -o gpurun_out/r01_syrkrows_c5 python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5a.log 2>&1
$NCU -k regex:k_trsm_ll -s 3 -c 1 -f -o gpurun_out/r01_trsm_c5 python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5b.log 2>&1
$NCU -k regex:k_search -s 1 -c 1 -f -o gpurun_out/r01_search_c5 python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5c.log 2>&1
tail -3 gpurun_out/ncu_c5*.log

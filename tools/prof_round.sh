set -x
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/r01_s3_bench_plain.json 2> gpurun_out/r01_s3_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r01_s3_c2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/prof_driver.py c2 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chol_small -s 5 -c 1 -o gpurun_out/r01_s3_chol_small_c2 -f python tools/prof_driver.py c2 3 > gpurun_out/ncu_c2chol.log 2>&1
RSLAM_C5_FILTERS=1024 python tools/prof_driver.py c5 2 > gpurun_out/plain_c5.log 2>&1 && \
RSLAM_C5_FILTERS=1024 ncu --set full --clock-control none --import-source on -k regex:k_syrk_rows -s 3 -c 1 -o gpurun_out/r01_s3_syrk_rows_c5 -f python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5a.log 2>&1
RSLAM_C5_FILTERS=1024 ncu --set full --clock-control none --import-source on -k regex:k_search -s 1 -c 1 -o gpurun_out/r01_s3_search_c5 -f python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5b.log 2>&1
RSLAM_C5_FILTERS=1024 ncu --set full --clock-control none --import-source on -k regex:k_trsm_small -s 3 -c 1 -o gpurun_out/r01_s3_trsm_small_c5 -f python tools/prof_driver.py c5 2 > gpurun_out/ncu_c5c.log 2>&1
ls -la gpurun_out/r01_s3_*

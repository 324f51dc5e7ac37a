# ncu evidence for profiles/ (run under gpurun; every ncu command follows a plain run of the same command that exited 0)
set -x
T=${1:-r01_s4}
python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/${T}_bench_plain.json 2> gpurun_out/${T}_bench_plain.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_c2_bench_launches.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/prof_driver.py c2 3 > gpurun_out/plain_c2.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_chol_small -s 5 -c 1 -o gpurun_out/${T}_chol_small_c2 -f python tools/prof_driver.py c2 3 > gpurun_out/ncu_c2chol.log 2>&1
RSLAM_DEDUPE=0 python tools/prof_driver.py c4 1 > gpurun_out/plain_c4b.log 2>&1 && \
RSLAM_DEDUPE=0 ncu --set full --clock-control none --import-source on -k regex:k_ransac_support -c 1 -o gpurun_out/${T}_support_c4_brute -f python tools/prof_driver.py c4 1 > gpurun_out/ncu_c4b.log 2>&1
python tools/prof_driver.py c4 1 > gpurun_out/plain_c4d.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_ransac_support -c 1 -o gpurun_out/${T}_support_c4_dedupe -f python tools/prof_driver.py c4 1 > gpurun_out/ncu_c4d.log 2>&1
true
true
ls -la gpurun_out/${T}_*

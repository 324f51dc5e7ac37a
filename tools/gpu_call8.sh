#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/t8.log 2>&1
tail -15 gpurun_out/t8.log
python tools/phase_clocks.py ransac_slam_b200/lib/librslam_dbg.so 2>&1 | tail -3
python bench.py --steps 200 --warmup 10 --extras c5,c3 --no-cpu > gpurun_out/b_8.json 2> gpurun_out/b_8.err
python tools/show_bench.py gpurun_out/b_8.json 2>/dev/null | grep -v "^  k_\(upd\|ransac\|ekf\|set\|pred\)"
tail -5 gpurun_out/b_8.err

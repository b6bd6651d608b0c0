#!/bin/bash
# final-build evidence: GPU test-suite, bench line, launch list of a bench step, full ncu capture of unpool_fused_kernel
cd /root/repo
timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r5_tests.txt
python bench.py --steps 5 --warmup 3 > gpurun_out/r5b_bench.json 2> gpurun_out/r5b_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r5_launches_raw.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/r5_launches_ncu.log 2>&1
python tools/unpool_bench.py --fused-only --points 5000 --iters 3 > gpurun_out/r5_unpool_5000.txt 2>&1
ncu --set full --clock-control none --import-source on -k regex:unpool_fused -s 1 -c 1 -o gpurun_out/r5_unpool_final -f python tools/unpool_bench.py --fused-only --points 5000 --iters 2 > gpurun_out/r5_unpool_final_ncu.log 2>&1
cat gpurun_out/r5_tests.txt; tail -c 600 gpurun_out/r5b_bench.json; cat gpurun_out/r5_unpool_5000.txt | tail -3

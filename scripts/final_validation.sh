#!/bin/bash
# final validation of the round: tests, bench, ncu evidence
timeout 300 python -m pytest tests -m gpu -q --tb=short 2>&1 | tail -3 > gpurun_out/final_tests.log
timeout 120 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline"
timeout 120 $CMD > gpurun_out/plain.log 2>&1 && timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_final.csv $CMD > gpurun_out/ncu1.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k k_scan -s 1 -c 1 -o gpurun_out/prof_scan_final -f $CMD > gpurun_out/ncu2.log 2>&1
cat gpurun_out/final_tests.log; tail -2 gpurun_out/ncu2.log

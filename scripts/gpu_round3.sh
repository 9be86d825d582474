#!/bin/bash
tag=${1:-r02c}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/gpu_tests_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_$tag.log
tail -4 gpurun_out/gpu_tests_$tag.log
timeout 600 python scripts/t_compare.py 30 280 > gpurun_out/t_compare_$tag.json 2> gpurun_out/t_compare_$tag.err; echo "t_compare rc=$?"; cat gpurun_out/t_compare_$tag.json; tail -3 gpurun_out/t_compare_$tag.err
timeout 600 python bench.py --no-mp --no-cpu --no-c4 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_$tag.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_t_$tag.csv \
  python scripts/t_compare.py 16 120 > gpurun_out/t_ncu_$tag.log 2>&1
echo "ncu rc=$?"

#!/bin/bash
# GPU session: all GPU tests, then the bench with the driver's arguments at N=1.
tag=${1:-r02b}
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/gpu_tests_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_$tag.log
tail -6 gpurun_out/gpu_tests_$tag.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"
tail -c 400 gpurun_out/bench_$tag.err

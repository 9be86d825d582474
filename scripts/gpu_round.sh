#!/bin/bash
# One GPU-box session: parity tests, the bench, and the per-kernel launch list of a short bench run.
# usage: scripts/gpu_round.sh <tag> [pytest args]
tag=${1:-r02}
shift
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi_$tag.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q "$@" > gpurun_out/gpu_tests_$tag.log 2>&1
echo "pytest rc=$?" >> gpurun_out/gpu_tests_$tag.log
tail -5 gpurun_out/gpu_tests_$tag.log
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err
echo "bench rc=$?"
tail -c 600 gpurun_out/bench_$tag.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$tag.csv \
  python bench.py --steps 2 --warmup 3 --no-cpu --no-mp --no-c4 --t-triples 4 > gpurun_out/bench_ncu_$tag.log 2>&1
echo "ncu rc=$?"

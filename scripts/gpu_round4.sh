#!/bin/bash
tag=${1:-r02d}
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off \
  -k regex:'dgemm_tma|tf32x3_gemm|t_energy_cp|pack_tau|ladder_unpack|split_tf32' -c 40 -f -o gpurun_out/ncu_$tag \
  python scripts/ncu_targets.py 40 300 all > gpurun_out/ncu_$tag.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_$tag.log
ncu -i gpurun_out/ncu_$tag.ncu-rep --page raw --csv > gpurun_out/ncu_${tag}_raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_$tag.ncu-rep --page details --csv > gpurun_out/ncu_${tag}_details.csv 2>/dev/null
ls -la gpurun_out/ncu_$tag*
rm -f gpurun_out/ncu_$tag.ncu-rep      # gpurun brings back at most 64 MiB: keep the CSV exports, not the report

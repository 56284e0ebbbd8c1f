#!/bin/bash
# one --set full capture of hist_scan2_kernel on the C4 workload (296,000 series = 148 x 2 CTAs x 1,000)
rm -f gpurun_out/r2f_full_c4.ncu-rep
timeout 500 ncu --set full --import-source on --clock-control none -k regex:hist_scan2 -c 1 -o gpurun_out/r2f_full_c4 \
  python bench.py --workload c4 --series 296000 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r2f_full_c4.log 2>&1
tail -2 gpurun_out/r2f_full_c4.log; ls -la gpurun_out/r2f_full_c4.ncu-rep

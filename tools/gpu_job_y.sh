#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --kernel-name-base demangled -k "regex:umma_gemm_kernel<.*64, .*4, .*1>" -c 8 -o gpurun_out/ncu_r02_wgrad python bench.py --workload train --batch 16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/ncu_y.log 2>&1; echo "rc=$?"

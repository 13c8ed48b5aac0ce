#!/bin/bash
# ncu --set full captures of every kernel class of the forward (B=16, V=100, default math mode), one launch each
mkdir -p gpurun_out
cap() {  # name regex skip
  timeout 400 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$2" -s $3 -c 1 \
    -o gpurun_out/ncu_r02_$1 python bench.py --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_o_$1.log 2>&1
  echo "$1 rc=$?"
}
cap upconv_gemm 'umma_gemm_kernel<256, 2, 0>' 1
cap flash 'flash_attn_kernel' 3
cap patchify 'patchify_umma_kernel' 1
cap ipp 'input_preprocess_ss_kernel' 1
cap vox_scatter 'vox_scatter_kernel' 2
cap vox_fill 'vox_fill_kernel' 2
cap up0_lowres 'umma_gemm_kernel<64, 4, 0>' 1
cap linear 'umma_gemm_kernel<256, 2, 1>' 40
cap geglu 'umma_gemm_kernel<256, 2, 4>' 9
cap rowmax 'umma_gemm_kernel<256, 2, 2>' 10
cap trans_gather 'trans_gather_kernel' 1
ls -la gpurun_out/*.ncu-rep

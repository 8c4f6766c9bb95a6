#!/bin/bash
# ncu evidence for profiles/ (run on the GPU box, one GPU):  bash tools/profile_run.sh r02
# 1. launch list of the bench command (gpu__time_duration.sum; cold-cache, serialised -> compare shares)
# 2. one --set full capture per hot kernel (NT=30000: ncu's save/restore cannot hold the 62 GB
#    environment cache of NT=60000; per-image behaviour is identical, byte counts scale linearly)
TAG=${1:-r02}
mkdir -p gpurun_out
B="python bench.py --steps 4 --warmup 3 --nt 30000 --no-cpu-baseline --sweep-avg 0"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 700 -c 700 --csv \
    --log-file gpurun_out/launches_${TAG}.csv $B > gpurun_out/ncu_ll.log 2>&1
cap() {  # name regex skip
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$2 -s $3 -c 1 -f \
      -o gpurun_out/prof_${TAG}_$1 $B > gpurun_out/ncu_$1.log 2>&1
}
cap oz_gemm oz_gemm_kernel 200   # past the ~95 <8,2> launches of init_envs: a <8,4> projection launch
cap oz_slice_rows oz_slice_rows_kernel 6
cap krgram2 krgram2_kernel 5
cap fat fat_kernel_t 20
cap jacobi_cluster jacobi_cluster_kernel 5
cap qr_block qr_block_kernel 3
ls -la gpurun_out/*.ncu-rep

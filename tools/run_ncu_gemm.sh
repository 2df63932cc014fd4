# ncu --set full of the back-transform / her2k DMMA GEMM launches in the middle of a solve
TAG=${1:-r1x}
export DWHMC_NGROUP=1
timeout 300 python tools/prof_diag.py 24 64 1 > gpurun_out/prof_plain_gemm_$TAG.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:zgemm -s 40 -c 3 -f -o gpurun_out/prof_gemm_$TAG python tools/prof_diag.py 24 64 1 > gpurun_out/ncu_gemm_$TAG.log 2>&1
tail -2 gpurun_out/ncu_gemm_$TAG.log

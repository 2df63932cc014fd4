# usage: tools/run_ncu_kernel.sh TAG KERNEL_REGEX [SKIP] [L] [B]   -- one ncu --set full capture of a kernel inside prof_diag
TAG=$1; KR=$2; SKIP=${3:-0}; L=${4:-24}; B=${5:-64}
timeout 300 python tools/prof_diag.py $L $B 1 > gpurun_out/prof_plain_$TAG.log 2>&1 && timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KR -s $SKIP -c 1 -f -o gpurun_out/prof_$TAG python tools/prof_diag.py $L $B 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log

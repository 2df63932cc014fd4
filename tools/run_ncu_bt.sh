export DWHMC_NGROUP=1
python tools/prof_diag.py 24 64 1 > gpurun_out/prof_plain5.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:zgemm -s 105 -c 3 -o gpurun_out/prof_bt_r1d python tools/prof_diag.py 24 64 1 > gpurun_out/ncu5.log 2>&1
tail -2 gpurun_out/ncu5.log

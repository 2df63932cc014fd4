export DWHMC_NGROUP=1
python tools/prof_diag.py 24 64 1 > gpurun_out/prof_plain6.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2600 --csv --log-file gpurun_out/launches_r1d.csv python tools/prof_diag.py 24 64 1 > gpurun_out/ncu6.log 2>&1
tail -2 gpurun_out/ncu6.log

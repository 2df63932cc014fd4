# launch list of one batched diagonalisation + force evaluation (groups serialised so the list is in column order)
# usage: tools/run_ncu_list.sh TAG
TAG=${1:-r1x}
export DWHMC_NGROUP=1
python tools/prof_diag.py 24 64 1 > gpurun_out/prof_plain_$TAG.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2700 --csv --log-file gpurun_out/launches_$TAG.csv python tools/prof_diag.py 24 64 1 > gpurun_out/ncu_$TAG.log 2>&1
tail -2 gpurun_out/ncu_$TAG.log

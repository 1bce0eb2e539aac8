set -x
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_v9.log 2>&1; tail -2 gpurun_out/smoke_v9.log
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_r1v9.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_launches_r1v9.log 2>&1
tail -c 300 gpurun_out/ncu_launches_r1v9.log
ncu --set full --clock-control none -k regex:'k_pyin_probs|k_pyin_backtrack' -s 4 -c 2 -o gpurun_out/full_r1v9_probs -f python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/ncu_full_r1v9.log 2>&1
tail -2 gpurun_out/ncu_full_r1v9.log

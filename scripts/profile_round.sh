#!/bin/bash
# Run on the GPU box (via gpurun): bench line, ncu launch list of the same command, one --set full capture
# of every kernel of the hot path.  Outputs land in gpurun_out/ and are summarised into profiles/ by
# scripts/ncu_summary.py / scripts/launch_list_summary.py on the build machine.
set -x
TAG=${1:-r2}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --cli-utts 0 > gpurun_out/ncu_launches_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_stft_mel|k_pyin_energy|k_pyin_cmnd|k_pyin_probs|k_pyin_viterbi|k_pyin_backtrack|k_align_prior|k_pitch_partials' \
    -s 9 -c 12 -o gpurun_out/full_${TAG} -f python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --cli-utts 0 \
    > gpurun_out/ncu_full_${TAG}.log 2>&1
tail -2 gpurun_out/ncu_full_${TAG}.log

python scripts/gpu_quick.py 4000 C2 > gpurun_out/s16_quick.log 2>&1
grep -h '"rep": 2' gpurun_out/s16_quick.log
python scripts/gpu_quick.py 1000 C4 > gpurun_out/s16_quick_c4.log 2>&1
grep -h '"rep": 2' gpurun_out/s16_quick_c4.log
(timeout 900 python -m pytest tests -m gpu -q -x -k "viterbi or pyin or config1 or config2 or config4" 2>&1 | tail -3) > gpurun_out/s16_tests.log
tail -1 gpurun_out/s16_tests.log
python scripts/viterbi_selfcheck.py 4000 C2 > gpurun_out/s16_selfcheck.json 2> gpurun_out/s16_selfcheck.err; cat gpurun_out/s16_selfcheck.json

python scripts/gpu_quick.py 4000 C2 > gpurun_out/s9_quick.log 2>&1
grep -h '"rep": 2' gpurun_out/s9_quick.log
python scripts/gpu_quick.py 1000 C4 > gpurun_out/s9_quick_c4.log 2>&1
grep -h '"rep": 2' gpurun_out/s9_quick_c4.log
(timeout 1200 python -m pytest tests -m gpu -q -x -k "logmel or fbank or config or featurizer or cli_cache or golden or properties" 2>&1 | tail -3) > gpurun_out/s9_tests.log
tail -1 gpurun_out/s9_tests.log

python scripts/gpu_quick.py 4000 C2 > gpurun_out/s30_quick.log 2>&1
grep -h '"rep": 2' gpurun_out/s30_quick.log
(timeout 900 python -m pytest tests -m gpu -q -x -k "pyin or config1 or config2 or config4 or golden" 2>&1 | tail -3) > gpurun_out/s30_tests.log
tail -1 gpurun_out/s30_tests.log

python scripts/gpu_quick.py 4000 C2 > gpurun_out/s10_quick.log 2>&1
grep -h '"rep": 2' gpurun_out/s10_quick.log
python scripts/gpu_quick.py 1000 C4 > gpurun_out/s10_quick_c4.log 2>&1
grep -h '"rep": 2' gpurun_out/s10_quick_c4.log

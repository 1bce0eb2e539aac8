set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2r_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2r_smoke.log 2>&1
bash scripts/profile_round.sh r2r
for w in C3 C4 C5; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2r_bench_$w.json 2> gpurun_out/r2r_bench_$w.err; done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2r_bench_reference.json 2> gpurun_out/r2r_bench_reference.err
tail -3 gpurun_out/r2r_tests.log

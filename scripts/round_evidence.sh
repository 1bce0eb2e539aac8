set -x
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -15) > gpurun_out/r2u_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u_smoke.log 2>&1
bash scripts/profile_round.sh r2u
for w in C3 C4 C5; do python bench.py --workload $w --steps 3 --warmup 3 > gpurun_out/r2u_bench_$w.json 2> gpurun_out/r2u_bench_$w.err; done
python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2u_bench_reference.json 2> gpurun_out/r2u_bench_reference.err
tail -3 gpurun_out/r2u_tests.log
python scripts/viterbi_selfcheck.py 13100 C2 > gpurun_out/r2u_selfcheck_c2.json 2> gpurun_out/r2u_selfcheck_c2.err
python scripts/viterbi_selfcheck.py 1500 C4 > gpurun_out/r2u_selfcheck_c4.json 2> gpurun_out/r2u_selfcheck_c4.err
python -m roar_b200.build -DROAR_VIT_STATS --out=build/libroar_sup_stats.so > /dev/null 2>&1 || true
ROAR_SUP_LIB=build/libroar_sup_stats.so python scripts/viterbi_step_stats.py C2 4000 gpurun_out/r2u_vitstats.json > gpurun_out/r2u_vitstats.log 2>&1

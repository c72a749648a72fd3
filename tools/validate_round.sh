set -x
timeout 400 python -m pytest tests -m gpu -x -q > gpurun_out/r1b_pytest_gpu.log 2>&1; tail -2 gpurun_out/r1b_pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r1b_smoke.log 2>&1; tail -3 gpurun_out/r1b_smoke.log
for w in c4 c3 c5 c2; do timeout 400 python bench.py --workload $w > gpurun_out/r1b_bench_$w.json 2> gpurun_out/r1b_bench_$w.err; cut -c1-200 gpurun_out/r1b_bench_$w.json; done
timeout 300 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r1b_bench_ref_c4.json 2>&1; cut -c1-300 gpurun_out/r1b_bench_ref_c4.json
timeout 600 bash tools/profile_round.sh r1b > gpurun_out/r1b_profile_round.log 2>&1; tail -3 gpurun_out/r1b_profile_round.log

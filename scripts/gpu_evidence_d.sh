#!/bin/bash
# last GPU call of round 2 (about 4 GPU-minutes left): the whole GPU suite on the rebuilt default library, the new
# Cayley-graph Tanner code through bench.py (BP and SSF), the default bench line
mkdir -p gpurun_out
timeout 120 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -3 gpurun_out/r2q_pytest_gpu.log
timeout 40 python bench.py --workload cayley_lp_4320 --steps 3 --warmup 3 --cpu-seconds 3 > gpurun_out/r2q_bench_cayley_lp_4320.json 2> gpurun_out/r2q_bench_cayley.err; echo "bench cayley bp rc=$?"
timeout 30 python bench.py --workload cayley_lp_4320 --decoder ssf --steps 3 --warmup 3 --cpu-seconds 3 > gpurun_out/r2q_bench_ssf_cayley_lp_4320.json 2>> gpurun_out/r2q_bench_cayley.err; echo "bench cayley ssf rc=$?"
timeout 50 python bench.py --cpu-seconds 4 > gpurun_out/r2q_bench_hgp_34_n1600.json 2> gpurun_out/r2q_bench.err; echo "bench default rc=$?"
cut -c1-400 gpurun_out/r2q_bench_cayley_lp_4320.json gpurun_out/r2q_bench_ssf_cayley_lp_4320.json gpurun_out/r2q_bench_hgp_34_n1600.json

#!/bin/bash
# BASELINE.json configurations at their stated sizes, on the GPU box (gpurun):  scripts/baseline_sweeps.sh <tag>
#   configs[1]  (3,4) HGP n=1600 and n=6400, BP (min-sum, 32 iterations) and SSF, p in geomspace(0.01, 0.10, 10), 1e7 shots per p
#   configs[2]  QC lifted product [[882,24]], BP, 1e8 shots (one GPU here; shots shard over ranks under torchrun)
#   configs[3]  PSL(2,5) lifted product n=2700, SSF sweep
# plus the LER-inside-the-oracle's-CI check (tests/tools/ler_ci_check.py).  CSVs go to gpurun_out/<tag>_*.
set -u
TAG=${1:-base}
OUT=gpurun_out
mkdir -p $OUT/codes
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from exp_ldpc_b200 import code_builders
from exp_ldpc_b200.codes import write_quantum_code
for name in ("pk19_b1_882", "a5_lift_2700"):
    with open(f"gpurun_out/codes/{name}.qecc", "w") as f:
        write_quantum_code(f, getattr(code_builders, name)())
PY
SW='(0.01,0.1,10)'
s() { name=$1; shift; python scripts/p_sweep.py "$@" > $OUT/${TAG}_sweep_${name}.csv 2>> $OUT/${TAG}_sweep.err; echo "sweep $name rc=$?"; }
s hgp1600_bp_ms32_1e7 codes/hgp_34_n1600.qecc --samples 10000000 --p_sweep "$SW" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s hgp1600_ssf_1e7 codes/hgp_34_n1600.qecc --samples 10000000 --p_sweep "$SW" --decoder_mode ssf --rounds 0
s hgp6400_bp_ms32_2e6 codes/hgp_34_n6400.qecc --samples 2000000 --p_sweep "$SW" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s hgp6400_ssf_1e6 codes/hgp_34_n6400.qecc --samples 1000000 --p_sweep "$SW" --decoder_mode ssf --rounds 0
s lp882_bp_ms32_1e8 $OUT/codes/pk19_b1_882.qecc --samples 100000000 --p_sweep "(0.05,0.05,1)" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s lp882_bp_ms32_sweep_1e7 $OUT/codes/pk19_b1_882.qecc --samples 10000000 --p_sweep "$SW" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s a5lp2700_ssf_1e6 $OUT/codes/a5_lift_2700.qecc --samples 1000000 --p_sweep "$SW" --decoder_mode ssf --rounds 0
s hgp1600_bposd_cs7_ms32_1e6 codes/hgp_34_n1600.qecc --samples 1000000 --p_sweep "$SW" --decoder_mode bposd --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
python tests/tools/ler_ci_check.py > $OUT/${TAG}_ler_ci_hgp1600_bp.csv 2>> $OUT/${TAG}_sweep.err; echo "ci bp rc=$?"
python tests/tools/ler_ci_check.py --decoder ssf --cpu-shots 20000 --same-shots 5000 > $OUT/${TAG}_ler_ci_hgp1600_ssf.csv 2>> $OUT/${TAG}_sweep.err; echo "ci ssf rc=$?"

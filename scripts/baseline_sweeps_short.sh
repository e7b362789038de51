#!/bin/bash
# Regression sweep after a kernel change (about two minutes of box time):  scripts/baseline_sweeps_short.sh <tag>
# The counts of the n=1600 BP sweep and of the [[882,24]] 1e8-shot point must equal the committed profiles/r2_sweep_* files
# (the shots are keyed by their global index and the fp64 path is bit-exact), whatever the kernel does internally.
set -u
TAG=${1:-base}
OUT=gpurun_out
mkdir -p $OUT/codes
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from exp_ldpc_b200 import code_builders
from exp_ldpc_b200.codes import write_quantum_code
with open("gpurun_out/codes/pk19_b1_882.qecc", "w") as f:
    write_quantum_code(f, code_builders.pk19_b1_882())
PY
SW='(0.01,0.1,10)'
s() { name=$1; shift; timeout 400 python scripts/p_sweep.py "$@" > $OUT/${TAG}_sweep_${name}.csv 2>> $OUT/${TAG}_sweep.err; echo "sweep $name rc=$?"; }
s hgp1600_bp_ms32_1e7 codes/hgp_34_n1600.qecc --samples 10000000 --p_sweep "$SW" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s lp882_bp_ms32_1e8 $OUT/codes/pk19_b1_882.qecc --samples 100000000 --p_sweep "(0.05,0.05,1)" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
# the reference's multi-round pipeline on the 1600-qubit code: storage circuit of 3 rounds under depolarizing_noise, sampled by
# Pauli frames, BP over the spacetime matrix (thread-block-cluster kernel), fold, logical flag
s hgp1600_rounds3_bp_2e5 codes/hgp_34_n1600.qecc --samples 200000 --p_sweep "(0.002,0.008,4)" --rounds 3 --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32
rm -rf $OUT/codes

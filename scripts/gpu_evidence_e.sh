#!/bin/bash
# p-sweeps of the Cayley-graph Tanner code (BASELINE configs[3]) with the reference's CLI: BP, SSF and BP+OSD-0
OUT=gpurun_out
mkdir -p $OUT/codes
python - <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from exp_ldpc_b200.cayley_codes import psl2_tanner_lp_4320
from exp_ldpc_b200.codes import write_quantum_code
with open("gpurun_out/codes/cayley_lp_4320.qecc", "w") as f:
    write_quantum_code(f, psl2_tanner_lp_4320())
PY
C=$OUT/codes/cayley_lp_4320.qecc
s() { name=$1; shift; timeout 32 python scripts/p_sweep.py $C "$@" > $OUT/r2q_sweep_cayley4320_${name}.csv 2>> $OUT/r2q_sweep.err; echo "sweep $name rc=$?"; }
s bp_ms32_1e6 --samples 1000000 --p_sweep "(0.002,0.03,8)" --decoder_mode bp --bposd_bp_method ms --bposd_max_iter 32 --rounds 0
s ssf_1e6 --samples 1000000 --p_sweep "(0.001,0.01,6)" --decoder_mode ssf --rounds 0
s bposd0_ms32_1e5 --samples 100000 --p_sweep "(0.005,0.03,5)" --decoder_mode bposd --bposd_bp_method ms --bposd_max_iter 32 --bposd_osd_method osd0 --rounds 0
rm -rf $OUT/codes
tail -n +1 $OUT/r2q_sweep_cayley4320_*.csv | cut -c1-110
tail -5 $OUT/r2q_sweep.err

#!/usr/bin/env python
"""Decode the shots of a GPU p-sweep CSV (scripts/p_sweep.py, code capacity, --decoder_mode bp | ssf | bposd) again on the CPU
oracle — same seed, same global shot indices, the flip probability and the prior computed exactly as p_sweep does — and
print both sets of counters side by side.  The fp64 BP path and SSF are bit-exact, so failures, non-converged shots
and the iteration total must be IDENTICAL.  Test infrastructure (it lives under tests/): it may import oracle/.

    python tests/tools/replicate_sweep_on_oracle.py <builder-or-qecc> <sweep.csv> [max_rows]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import pandas as pd  # noqa: E402

from exp_ldpc_b200 import cayley_codes, code_builders  # noqa: E402
from exp_ldpc_b200.codes import load_code  # noqa: E402
from exp_ldpc_b200.storage import depolarizing_noise  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

src, csv = sys.argv[1], sys.argv[2]
rows = int(sys.argv[3]) if len(sys.argv) > 3 else None
if os.path.exists(src):
    code = load_code(src)
else:
    code = getattr(code_builders if hasattr(code_builders, src) else cayley_codes, src)()
hz, lz, hx = code.checks.z, code.logicals.z, code.checks.x
df = pd.read_csv(csv, index_col=0, float_precision="round_trip")     # p must come back to the last bit
print("decoder,p,shots,gpu_failures,cpu_failures,gpu_non_converged,cpu_non_converged,gpu_iters_total,cpu_iters_total,identical,cpu_seconds")
for _, r in df.iloc[:rows].iterrows():
    p = float(r.p_ph)
    q = depolarizing_noise(p, p).code_capacity_flip_probability()       # scripts/p_sweep.py: pm = p
    prior = 2 / 3 * p                                                   # constant_pheno_prior
    dec = r.decoder_mode
    oracle_dec = {"bposd": {"osd0": "bposd0", "osd_e": "bposd_e", "osd_cs": "bposd_cs"}[r.osd_method]}.get(dec, dec)
    t = time.time()
    c = po.run_code_capacity(hz, lz, q, prior, seed=0, shots=int(r.samples), decoder=oracle_dec, hx=hx,
                             max_iter=0 if dec == "ssf" else int(r.max_iter), bp_method=r.bp_method,
                             osd_order=0 if r.osd_method == "osd0" else int(r.osd_order),
                             ms_scaling_factor=float(r.ms_scaling_factor))
    same = (c["failures"], c["non_converged"], c["iters_total"]) == (int(r.failures), int(r.non_converged), int(r.iters_total))
    print(f"{dec},{p!r},{int(r.samples)},{int(r.failures)},{c['failures']},{int(r.non_converged)},{c['non_converged']},"
          f"{int(r.iters_total)},{c['iters_total']},{int(same)},{time.time() - t:.1f}", flush=True)

#!/usr/bin/env python
"""How often does the fp32 BP path reproduce the fp64 (reference-order) hard decisions?"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import probabilities  # noqa: E402
from exp_ldpc_b200.codes import load_code  # noqa: E402
from exp_ldpc_b200.experiment import CodeCapacityExperiment  # noqa: E402

shots = int(sys.argv[1]) if len(sys.argv) > 1 else 50000
for name in ("hgp_34_n225", "hgp_34_n1600"):
    code = load_code(os.path.join(ROOT, "codes", name + ".qecc"))
    for method, alpha in (("ms", 0.0), ("ms", 0.625), ("ps", 0.0)):
        for p in (0.01, 0.02, 0.05):
            q, prior = probabilities(p)
            opts = dict(max_iter=32, bp_method=method, ms_scaling_factor=alpha)
            a = CodeCapacityExperiment(code, q, prior, "bp", opts, precision="f64").run(shots, seed=1, per_shot=True)
            b = CodeCapacityExperiment(code, q, prior, "bp", opts, precision="f32").run(shots, seed=1, per_shot=True)
            same = (a["decodings"] == b["decodings"]).all(axis=1)
            print(f"{name} {method} alpha={alpha} p={p}: decisions agree {same.mean():.5%}  iters agree "
                  f"{(a['iters'] == b['iters']).mean():.5%}  LER f64 {a['failures'] / shots:.5f} f32 {b['failures'] / shots:.5f} "
                  f"nonconv f64 {a['non_converged']} f32 {b['non_converged']}", flush=True)

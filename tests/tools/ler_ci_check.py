#!/usr/bin/env python
"""North-star parity bar for BP (BASELINE.json): at every swept p the GPU decoder's logical error rate lies
inside the 95 % binomial (Wilson) interval of the CPU oracle's, measured on INDEPENDENT shots (different seed),
and — fp64 path — the hard decisions on identical shots are bit-identical (fp32: agreement fraction printed).

    python tests/tools/ler_ci_check.py [--code codes/hgp_34_n1600.qecc] [--gpu-shots 1000000] [--cpu-shots 50000]
Prints one CSV row per (precision, p).  Test infrastructure (it lives under tests/): it may import oracle/."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402

from bench import probabilities  # noqa: E402
from exp_ldpc_b200.codes import load_code  # noqa: E402
from exp_ldpc_b200.experiment import CodeCapacityExperiment, binomial_ci  # noqa: E402
from oracle import pyoracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--code", default=os.path.join(ROOT, "codes", "hgp_34_n1600.qecc"))
ap.add_argument("--gpu-shots", type=int, default=1000000)
ap.add_argument("--cpu-shots", type=int, default=50000)
ap.add_argument("--same-shots", type=int, default=20000, help="identical shots decoded by both for the decision agreement")
ap.add_argument("--p-sweep", default="0.01,0.1,10")
ap.add_argument("--max-iter", type=int, default=32)
ap.add_argument("--bp-method", default="ms")
ap.add_argument("--decoder", default="bp", choices=["bp", "ssf"])
a = ap.parse_args()
lo, hi, k = a.p_sweep.split(",")
code = load_code(a.code)
hz, lz = code.checks.z, code.logicals.z
print("precision,p,gpu_failures,gpu_shots,gpu_ler,cpu_failures,cpu_shots,cpu_ler,cpu_ci_lo,cpu_ci_hi,inside_ci,"
      "decisions_equal_fraction,same_shots")
for precision in (("f64", "f32") if a.decoder == "bp" else ("f64",)):
    for p in np.geomspace(float(lo), float(hi), int(k)):
        q, prior = probabilities(p)
        opts = dict(max_iter=a.max_iter, bp_method=a.bp_method, ms_scaling_factor=0.0)
        exp = CodeCapacityExperiment(code, q, prior, a.decoder, opts, precision=precision)
        g = exp.run(a.gpu_shots, seed=0)
        # (the oracle's SSF takes its round cap from max_iter: 0 = no cap, like the GPU decoder's default)
        kw = dict(decoder=a.decoder, hx=code.checks.x, max_iter=a.max_iter if a.decoder == "bp" else 0, bp_method=a.bp_method,
                  ms_scaling_factor=0.0)
        c = pyoracle.run_code_capacity(hz, lz, q, prior, seed=12345, shots=a.cpu_shots, **kw)
        ci = binomial_ci(c["failures"], c["shots"])
        gl = g["failures"] / g["shots"]
        gs = exp.run(a.same_shots, seed=7, per_shot=True)
        cs = pyoracle.run_code_capacity(hz, lz, q, prior, seed=7, shots=a.same_shots, per_shot=True, **kw)
        same = float((gs["decodings"] == cs["decodings"]).all(axis=1).mean())
        print(f"{precision},{p:.6g},{g['failures']},{g['shots']},{gl:.6g},{c['failures']},{c['shots']},"
              f"{c['failures'] / c['shots']:.6g},{ci[0]:.6g},{ci[1]:.6g},{int(ci[0] - 1e-12 <= gl <= ci[1] + 1e-12)},{same:.6f},{a.same_shots}",
              flush=True)

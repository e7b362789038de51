#!/usr/bin/env python
"""Write tests/golden/decoder_vectors.npz: hard decisions / iteration counts / failure flags of
the CPU oracle on the fixed-seed cases of tests/tests_golden_cases.py.  These are OUR oracle's
outputs (the reference has no decoder vectors and `ldpc` is absent: "parity unpinned")."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

from exp_ldpc_b200.codes import load_code  # noqa: E402
from tests_golden_cases import CASES, run_case  # noqa: E402

codes = {"hgp225": load_code(os.path.join(ROOT, "codes", "hgp_34_n225.qecc")),
         "hgp1600": load_code(os.path.join(ROOT, "codes", "hgp_34_n1600.qecc"))}
out = {}
for name in CASES:
    r = run_case(name, codes)
    out[name + "_dec"] = np.packbits(r["decodings"], axis=1)
    out[name + "_iters"] = r["iters"]
    out[name + "_fail"] = r["fail"]
    out[name + "_conv"] = r["converged"]
    print(name, "failures", r["failures"], "non_converged", r["non_converged"], "iters", r["iters_total"])
np.savez_compressed(os.path.join(HERE, "decoder_vectors.npz"), **out)

#!/usr/bin/env python
"""Stress the cluster mode of the thread-owned-state BP kernel on a small graph with many short shots (every cluster runs
into the end of the shot queue): decode one golden case K times with a forced (cluster size, slots) geometry and compare
every run with the committed vectors.  Usage: python scripts/stress_cluster.py <case> <ctas> <slots> <repeats>
Run it under `timeout`: a hang is the failure this looks for."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import numpy as np  # noqa: E402


def main(case, ctas, slots, repeats):
    from exp_ldpc_b200.codes import load_code
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from tests_golden_cases import CASES, case_probabilities
    key, decoder, opts, p, shots, seed = CASES[case]
    code = load_code(os.path.join(ROOT, "codes", {"hgp225": "hgp_34_n225.qecc", "hgp1600": "hgp_34_n1600.qecc"}[key]))
    g = np.load(os.path.join(ROOT, "tests", "golden", "decoder_vectors.npz"))
    q, prior = case_probabilities(p)
    exp = CodeCapacityExperiment(code, q, prior, decoder, opts, kernel="regular_cluster")
    exp.decoder.configure(ctas_per_sm=ctas, slots=slots)
    print(case, exp.decoder.config(), flush=True)
    for i in range(repeats):
        out = exp.run(shots, seed=seed, per_shot=True)
        assert np.array_equal(np.packbits(out["decodings"], axis=1), g[case + "_dec"]), i
        assert np.array_equal(out["iters"], g[case + "_iters"]), i
        if i % 20 == 19:
            print("  ", i + 1, "runs ok", flush=True)
    print("done", flush=True)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))

#!/usr/bin/env python
"""Detector-error-model decoding (decoder_mode 'bpd_detector', _experiment.py:128-151) of the [[225,9]] code under
circuit_noise: the model of T rounds has ~15 k faults / ~96 k edges at T=2 and check degrees above 200, so BP runs on the
global-memory kernel (bpg_kernel.cuh).  Prints the model's size and the decoded shots/s (sampler + detectors + BP + flag).
Usage: python scripts/microbench/dem_bp.py [rounds] [shots]"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402


def main(rounds=2, shots=16384):
    import torch
    from exp_ldpc_b200.codes import load_code
    from exp_ldpc_b200.experiment import DetectorExperiment
    from exp_ldpc_b200.storage import build_storage_simulation, circuit_noise
    code = load_code(os.path.join(ROOT, "codes", "hgp_34_n225.qecc"))
    opts = dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0)
    exp = DetectorExperiment(build_storage_simulation(rounds, circuit_noise(0.002, 0.002), code).circuit, opts)
    h = exp.dem.fault_check_matrix.tocsr()
    col_deg = np.diff(h.tocsc().indptr)
    row_deg = np.diff(h.indptr)
    print(f"model: {h.shape[0]} detectors x {h.shape[1]} faults, {h.nnz} edges, column degree max {col_deg.max()} mean {col_deg.mean():.2f}, "
          f"row degree max {row_deg.max()} mean {row_deg.mean():.1f}; kernel {exp.decoder.config()}")
    exp.run(min(shots, 2048), seed=1)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = exp.run(shots, seed=2)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f"rounds={rounds} shots={shots}: {shots / dt:.0f} shots/s ({dt * 1e3:.1f} ms), failures {out['failures']}, "
          f"non-converged {out['non_converged']}, iterations/shot {out['iters_total'] / shots:.2f}")


if __name__ == "__main__":
    main(*(int(x) for x in sys.argv[1:3]))

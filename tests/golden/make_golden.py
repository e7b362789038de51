#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ and the pinned benchmark codes under codes/.

Runs ONLY in the build container: it imports the read-only reference (``/root/reference/python``)
with import-only stand-ins for ``galois``/``stim`` (``tests/golden/_stubs``), calls the reference's
own generators / builders, and stores their outputs.  The GPU box never runs this (``/root/reference``
does not exist there); tests read the committed files.

What comes from the reference itself (so the oracle and the host mirror are *pinned* on it):
  * Hx/Hz of the (3,4) HGP codes  — ``qldpc.biregular_hgp`` (``hypergraph_product_code.py:7-35``)
  * qecc text                     — ``qldpc.write_quantum_code`` / ``read_quantum_code``
  * spacetime matrices, syndromes, folds — ``qldpc.SpacetimeCode*`` (``spacetime_code.py``)
  * measurement-record layout + circuit text — ``qldpc.build_storage_simulation`` (``storage_sim.py``)
  * noise-rewrite output          — ``qldpc.noise_model`` (``noise_model.py``)
What does not: logical operators (galois is absent) — computed by ``exp_ldpc_b200.gf2`` and then
validated through the reference's ``read_quantum_code(validate_stabilizer_code=True)``.

Usage:  python tests/golden/make_golden.py
"""
import gzip
import hashlib
import io
import json
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [os.path.join(HERE, "_stubs"), "/root/reference/python", ROOT]
warnings.filterwarnings("ignore")

import numpy as np  # noqa: E402
from scipy import sparse  # noqa: E402

import qldpc  # noqa: E402  (the reference)
import qldpc.spacetime_code as ref_st  # noqa: E402

from exp_ldpc_b200 import gf2  # noqa: E402
from exp_ldpc_b200.codes import QuantumCode, QuantumCodeChecks, QuantumCodeLogicals, write_quantum_code  # noqa: E402

# scipy >= 1.14 rejects the iterator the reference passes to block_diag (spacetime_code.py:52);
# hand it a list instead, without touching the reference.
ref_st.repeat = lambda x, n: [x] * n

CODES = os.path.join(ROOT, "codes")
os.makedirs(CODES, exist_ok=True)


def csr_triplet(m):
    m = sparse.csr_matrix(m)
    m.sum_duplicates()
    m.sort_indices()
    return m.indptr.astype(np.int32), m.indices.astype(np.int32), np.asarray(m.shape, dtype=np.int32)


def make_hgp(name, nv, seed, girth_bound=None):
    ref = qldpc.biregular_hgp(nv, 3, 4, seed=seed, compute_logicals=False, girth_bound=girth_bound)
    lx, lz = gf2.css_logicals(ref.checks.x, ref.checks.z)
    code = QuantumCode(QuantumCodeChecks(ref.checks.x, ref.checks.z), QuantumCodeLogicals(lx, lz))
    buf = io.StringIO()
    buf.write(f"c (3,4) hypergraph product code, qldpc.biregular_hgp({nv},3,4,seed={seed},girth_bound={girth_bound})\n")
    buf.write("c checks: reference generator; logicals: exp_ldpc_b200.gf2.css_logicals\n")
    write_quantum_code(buf, code)
    text = buf.getvalue()
    # the reference's reader must accept it, validation on, and give back the same matrices
    back = qldpc.read_quantum_code(io.StringIO(text), validate_stabilizer_code=True)
    assert (back.checks.x != ref.checks.x).nnz == 0 and (back.checks.z != ref.checks.z).nnz == 0
    assert np.array_equal(back.logicals.z, lz) and np.array_equal(back.logicals.x, lx)
    path = os.path.join(CODES, name + ".qecc")
    with open(path, "w") as fh:
        fh.write(text)
    print(name, code.num_qubits, code.checks.z.shape, "k =", code.num_logicals)
    return ref, code, text


def main():
    meta = {}
    ref225, code225, text225 = make_hgp("hgp_34_n225", 12, 0)
    make_hgp("hgp_34_n1600", 32, 0)
    make_hgp("hgp_34_n6400", 64, 0)
    make_hgp("hgp_34_n2025_test", 36, 42, girth_bound=4)  # random_test_hgp(), hypergraph_product_code.py:37-40

    # ---- qecc text written by the REFERENCE writer (reader parity fixture) -----------------
    buf = io.StringIO()
    ref_code = qldpc.QuantumCode(ref225.checks, qldpc.QuantumCodeLogicals(
        np.array(code225.logicals.x, dtype=np.uint32), np.array(code225.logicals.z, dtype=np.uint32)))
    qldpc.write_quantum_code(buf, ref_code)
    with open(os.path.join(HERE, "hgp_34_n225_refwriter.qecc"), "w") as fh:
        fh.write(buf.getvalue())

    hz = ref225.checks.z
    m, n = hz.shape
    rng = np.random.default_rng(1234)
    out = {}
    ip, ix, shp = csr_triplet(hz)
    out["hz_indptr"], out["hz_indices"], out["hz_shape"] = ip, ix, shp

    # ---- spacetime matrices / syndrome / fold from the reference ----------------------------
    for T in (0, 1, 3):
        st = qldpc.SpacetimeCode(hz, T)
        ip, ix, shp = csr_triplet(st.spacetime_check_matrix)
        out[f"st{T}_indptr"], out[f"st{T}_indices"], out[f"st{T}_shape"] = ip, ix, shp
        hist = rng.integers(0, 2, size=(max(T, 1), m)).astype(np.int64)
        readout = rng.integers(0, 2, size=n).astype(np.int64)
        synd = st.syndrome_from_history(lambda t, hist=hist: hist[t], readout)
        fault = rng.integers(0, 2, size=st.spacetime_check_matrix.shape[1]).astype(np.int64)
        out[f"st{T}_hist"], out[f"st{T}_readout"], out[f"st{T}_synd"] = hist, readout, np.asarray(synd)
        out[f"st{T}_fault"], out[f"st{T}_fold"] = fault, np.asarray(st.final_correction(fault))
        prior = np.zeros(st.spacetime_check_matrix.shape[1])
        st.data_bits(prior)[:] = 0.25
        st.measurement_bits(prior)[:] = 0.5
        out[f"st{T}_prior_quirk"] = prior
    ss = qldpc.SpacetimeCodeSingleShot(hz)
    ip, ix, shp = csr_triplet(ss.spacetime_check_matrix)
    out["ss_indptr"], out["ss_indices"], out["ss_shape"] = ip, ix, shp
    np.savez_compressed(os.path.join(HERE, "spacetime_n225.npz"), **out)

    # ---- storage-simulation record layout + circuit text ------------------------------------
    layout = {}
    qcode = qldpc.read_quantum_code(io.StringIO(text225))
    for T in (0, 1, 3):
        sim = qldpc.build_storage_simulation(T, qldpc.noise_model.depolarizing_noise(0.05, 0.05), qcode,
                                             use_x_logicals=False)
        circuit = "\n".join(sim.circuit)
        n_meas = T * (qcode.checks.x.shape[0] + m) + n
        rec = np.arange(n_meas)
        layout[str(T)] = {
            "record_length": int(n_meas),
            "data_view": [int(sim.data_view(rec)[0]), int(sim.data_view(rec)[-1])],
            "z_views": [[int(sim.measurement_view(t, False, rec)[0]), int(sim.measurement_view(t, False, rec)[-1])]
                        for t in range(T)],
            "x_views": [[int(sim.measurement_view(t, True, rec)[0]), int(sim.measurement_view(t, True, rec)[-1])]
                        for t in range(T)],
            "circuit_sha256": hashlib.sha256(circuit.encode()).hexdigest(),
            "circuit_lines": len(sim.circuit),
        }
        with gzip.open(os.path.join(HERE, f"storage_sim_n225_T{T}_depol.stim.gz"), "wt") as fh:
            fh.write(circuit)
    sim = qldpc.build_storage_simulation(1, qldpc.noise_model.circuit_noise(0.01, 0.02), qcode, use_x_logicals=False)
    with gzip.open(os.path.join(HERE, "storage_sim_n225_T1_circuit.stim.gz"), "wt") as fh:
        fh.write("\n".join(sim.circuit))
    meta["storage_layout"] = layout

    # ---- noise-rewrite goldens (the reference's own test vectors re-derived from its code) ----
    toy = ['RX 0 1 2', 'TICK', 'CZ 0 1', 'TICK', 'MX 0 2', 'TICK', 'TICK', 'MX 0']
    targets = qldpc.CircuitTargets([1], [0, 2], [])
    meta["noise_pheno"] = list(qldpc.noise_model.depolarizing_noise(0.1, 0.2).rewrite(targets, toy))
    meta["noise_circuit"] = list(qldpc.noise_model.circuit_noise(0.1, 0.2).rewrite(targets, toy))

    with open(os.path.join(HERE, "reference_meta.json"), "w") as fh:
        json.dump(meta, fh, indent=1, sort_keys=True)
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()

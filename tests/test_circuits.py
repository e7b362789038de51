"""Circuit-level sampling (CPU side): the Stim-text compiler on the reference-generated circuit
fixtures, and the oracle Pauli-frame sampler's physics."""
import gzip
import os

import numpy as np
import pytest

from conftest import GOLDEN
from exp_ldpc_b200.circuits import compile_circuit


def _golden(name):
    with gzip.open(os.path.join(GOLDEN, f"storage_sim_n225_{name}.stim.gz"), "rt") as fh:
        return fh.read().split("\n")


@pytest.mark.parametrize("name,rounds", [("T0_depol", 0), ("T1_depol", 1), ("T3_depol", 3), ("T1_circuit", 1)])
def test_compile_reference_circuits(name, rounds):
    """Measurement count = T(mx+mz)+n (storage_sim.py:187-196; tests/test_storage_sim.py:107-136)."""
    prog = compile_circuit(_golden(name))
    assert prog.num_measurements == rounds * 216 + 225
    assert prog.num_qubits == (225 if rounds == 0 else 441)
    assert prog.thresholds[0] == 0 and prog.ops.dtype == np.int32
    sites = prog.ops[(prog.ops[:, 0] >> 8) > 0, 3]
    assert np.array_equal(sites, np.arange(prog.num_sites))               # every noisy op has its own RNG site


def test_repeat_blocks_unroll():
    a = compile_circuit(["RZ 0 1", "REPEAT 3 {", "CX 0 1", "M(0.5) 1", "}", "M 0"])
    b = compile_circuit(["RZ 0 1"] + ["CX 0 1", "M(0.5) 1"] * 3 + ["M 0"])
    assert np.array_equal(a.ops, b.ops) and a.num_measurements == 4 and a.num_sites == 3
    with pytest.raises(RuntimeError):
        compile_circuit(["H 0"])
    with pytest.raises(RuntimeError):
        compile_circuit(["REPEAT 2 {", "M 0"])


def test_frame_propagation_rules():
    """Deterministic frame checks with p = 1 channels."""
    from oracle import pyoracle as po

    def rec(lines):
        p = compile_circuit(lines)
        return po.circuit_sample(p.ops, p.thresholds, p.num_qubits, p.num_measurements, 0, 0, 1)[0]
    # X on the control of a CX spreads to the target; Z on the target spreads to the control
    assert list(rec(["RZ 0 1", "X_ERROR(1) 0", "CX 0 1", "M 0 1"])) == [1, 1]
    assert list(rec(["RX 0 1", "Z_ERROR(1) 1", "CX 0 1", "MX 0 1"])) == [1, 1]
    # CZ: X on one side becomes Z on the other -> flips an X-basis measurement there
    assert list(rec(["RX 0", "RZ 1", "X_ERROR(1) 1", "CZ 0 1", "MX 0", "M 1"])) == [1, 1]
    # Z errors do not flip Z-basis measurements; measure-and-reset clears the frame
    assert list(rec(["RZ 0", "Z_ERROR(1) 0", "M 0"])) == [0]
    assert list(rec(["RX 0", "Z_ERROR(1) 0", "MRX 0", "MX 0"])) == [1, 0]
    # measurement flip probability 1
    assert list(rec(["RZ 0", "M(1) 0"])) == [1]


def test_noiseless_circuits_give_zero_records():
    from oracle import pyoracle as po
    for name in ("T1_depol", "T1_circuit"):
        p = compile_circuit(_golden(name))
        r = po.circuit_sample(p.ops, np.zeros_like(p.thresholds), p.num_qubits, p.num_measurements, 3, 0, 20)
        assert not r.any()


def test_code_capacity_circuit_matches_closed_form():
    """rounds = 0 under depolarizing_noise(p, p): data flip rate = 2p/3 + p - 4p^2/3 — the rate the
    code-capacity sampler uses (noise_model.py:117-123; exp_ldpc_b200.experiment.NoiseModel)."""
    from exp_ldpc_b200.experiment import depolarizing_noise
    from oracle import pyoracle as po
    p = compile_circuit(_golden("T0_depol"))
    r = po.circuit_sample(p.ops, p.thresholds, p.num_qubits, p.num_measurements, 5, 0, 4000)
    q = depolarizing_noise(0.05, 0.05).code_capacity_flip_probability()
    assert abs(r.mean() - q) < 4 * np.sqrt(q * (1 - q) / r.size)


def test_syndrome_statistics_of_one_round():
    """T = 1, depolarizing_noise(0.05, 0.05): X-check bits flip with pm only; a Z check of weight 7
    sees the parity of 7 data X/Y errors (2p/3 each), then the readout flip."""
    from oracle import pyoracle as po
    p = compile_circuit(_golden("T1_depol"))
    r = po.circuit_sample(p.ops, p.thresholds, p.num_qubits, p.num_measurements, 6, 0, 3000)
    a = 2 * 0.05 / 3
    par = (1 - (1 - 2 * a) ** 7) / 2
    zexp = par * 0.95 + (1 - par) * 0.05
    assert abs(r[:, :108].mean() - 0.05) < 0.003
    assert abs(r[:, 108:216].mean() - zexp) < 0.004

"""Host mirror of the storage-simulation circuit builder and noise models (CPU).

Pinned on the reference's own golden vectors (tests/test_storage_sim.py:13-77 via
tests/golden/reference_meta.json), its structural tests (tests/test_storage_sim.py:79-154,
tests/test_edge_coloring.py) and the reference-generated circuit fixtures."""
import gzip
import json
import os

import numpy as np
import pytest

from conftest import GOLDEN
from exp_ldpc_b200 import storage
from exp_ldpc_b200.circuits import compile_circuit

SMALL = ['RX 0 1 2', 'TICK', 'CZ 0 1', 'TICK', 'MX 0 2', 'TICK', 'TICK', 'MX 0']


def _meta():
    return json.load(open(os.path.join(GOLDEN, "reference_meta.json")))


def _golden(name):
    with gzip.open(os.path.join(GOLDEN, f"storage_sim_n225_{name}.stim.gz"), "rt") as fh:
        return fh.read().split("\n")


def test_noise_rewrite_goldens():
    """tests/test_storage_sim.py:13-77 of the reference (outputs recorded by running the reference)."""
    t = storage.CircuitTargets([1], [0, 2], [])
    assert storage.depolarizing_noise(0.1, 0.2).rewrite(t, SMALL) == _meta()["noise_pheno"]
    assert storage.circuit_noise(0.1, 0.2).rewrite(t, SMALL) == _meta()["noise_circuit"]
    assert storage.trivial_noise().rewrite(t, SMALL) == SMALL


@pytest.mark.parametrize("seed", range(6))
def test_schedule_is_a_proper_colouring(seed):
    """tests/test_edge_coloring.py:6-46: #colours == max degree, no vertex twice in a colour."""
    from scipy import sparse
    rng = np.random.default_rng(seed)
    h = sparse.csr_matrix((rng.random((14, 23)) < 0.25).astype(np.uint32))
    layers = storage.schedule_checks(h)
    deg = int(max(h.sum(axis=0).max(), h.sum(axis=1).max()))
    assert len(layers) == deg
    seen = set()
    for layer in layers:
        assert len(set(layer.values())) == len(layer)              # a data qubit once per layer (checks: dict keys)
        for c, q in layer.items():
            assert h[c, q] == 1 and (c, q) not in seen
            seen.add((c, q))
    assert len(seen) == h.nnz


def _check_colouring(a, colours):
    """tests/test_edge_coloring.py:6-18: max-degree many colours, every edge once, no vertex twice in a colour."""
    a = np.asarray(a)
    assert len(colours) == max(a.sum(axis=0).max(), a.sum(axis=1).max())
    edges = [e for c in colours for e in c]
    assert len(edges) == len(set(edges)) == a.sum()
    for u in range(a.shape[0]):
        for v in range(a.shape[1]):
            assert sorted(k for (x, y, k) in edges if (x, y) == (u, v)) == list(range(a[u, v]))
    for c in colours:
        assert len({u for u, _, _ in c}) == len(c) and len({v for _, v, _ in c}) == len(c)


@pytest.mark.parametrize("seed", range(12))
def test_bipartite_edge_colouring(seed):                           # tests/test_edge_coloring.py:21-28
    rng = np.random.default_rng(seed)
    nl, nr, degree = int(rng.integers(10, 40)), int(rng.integers(10, 40)), int(rng.integers(4, 12))
    a = (rng.random((nl, nr)) < degree / np.sqrt(nl * nr)).astype(np.int64)
    _check_colouring(a, storage.edge_color_bipartite(a))


@pytest.mark.parametrize("seed", range(12))
def test_bipartite_multigraph_edge_colouring(seed):                # tests/test_edge_coloring.py:30-46
    rng = np.random.default_rng(100 + seed)
    nl, nr = int(rng.integers(6, 25)), int(rng.integers(6, 25))
    a = np.zeros((nl, nr), dtype=np.int64)
    for u in range(nl):                                            # configuration model: stubs of u land anywhere, repeats allowed
        for v in rng.integers(0, nr, size=int(rng.integers(4, 12))):
            a[u, v] += 1
    assert a.max() >= 2                                            # parallel edges are present
    colours = storage.edge_color_bipartite(a)
    _check_colouring(a, colours)
    from scipy import sparse
    assert storage.edge_color_bipartite(sparse.csr_matrix(a)) == colours


def test_edge_colouring_degenerate_inputs():
    assert storage.edge_color_bipartite(np.zeros((3, 4), dtype=int)) == []
    assert storage.edge_color_bipartite(np.array([[3]])) == [{(0, 0, 0)}, {(0, 0, 1)}, {(0, 0, 2)}] or \
        sorted(map(sorted, storage.edge_color_bipartite(np.array([[3]])))) == [[(0, 0, 0)], [(0, 0, 1)], [(0, 0, 2)]]
    with pytest.raises(RuntimeError):
        storage.edge_color_bipartite(np.array([[1, -1]]))


def test_ancilla_targets_reconstruct_the_checks(code225):
    """tests/test_storage_sim.py:79-105."""
    targets, circuit = storage.build_perfect_circuit(code225)
    n, mx = 225, 108
    hx = np.zeros((108, 225), int)
    hz = np.zeros((108, 225), int)
    for ln in circuit:
        tk = ln.split()
        if tk[0] == "CX":
            hx[int(tk[1]) - n, int(tk[2])] += 1
        elif tk[0] == "CZ":
            hz[int(tk[1]) - n - mx, int(tk[2])] += 1
    assert np.array_equal(hx, code225.checks.x.toarray()) and np.array_equal(hz, code225.checks.z.toarray())
    assert targets.data == list(range(225)) and targets.x_checks[0] == 225 and targets.z_checks[0] == 333
    for step in "\n".join(circuit).split("TICK"):                  # a qubit appears once per time step (storage_sim.py:89-108)
        q = [int(x) for x in step.split() if x.isdigit()]
        assert len(q) == len(set(q))


@pytest.mark.parametrize("rounds", [0, 1, 3])
@pytest.mark.parametrize("use_x", [False, True])
def test_record_layout_and_views(code225, rounds, use_x):
    """tests/test_storage_sim.py:107-154 and the layout recorded from the reference."""
    sim = storage.build_storage_simulation(rounds, storage.depolarizing_noise(0.1, 0), code225, use_x_logicals=use_x)
    prog = compile_circuit(sim.circuit)
    assert prog.num_measurements == rounds * 216 + 225
    rec = np.arange(prog.num_measurements)
    assert len(sim.data_view(rec)) == 225 and sim.data_view(rec).base is rec           # a view, not a copy
    for r in range(rounds):
        for gx in (True, False):
            v = sim.measurement_view(r, gx, rec)
            assert len(v) == 108 and v.base is rec
    if not use_x:
        lay = _meta()["storage_layout"][str(rounds)]
        assert lay["record_length"] == prog.num_measurements
        assert [int(sim.data_view(rec)[0]), int(sim.data_view(rec)[-1])] == lay["data_view"]
        for r in range(rounds):
            assert [int(sim.measurement_view(r, False, rec)[0]), int(sim.measurement_view(r, False, rec)[-1])] == lay["z_views"][r]
            assert [int(sim.measurement_view(r, True, rec)[0]), int(sim.measurement_view(r, True, rec)[-1])] == lay["x_views"][r]
    ndet = sum(ln.startswith("DETECTOR") for ln in "\n".join(sim.circuit).split("\n"))
    nobs = sum(ln.startswith("OBSERVABLE_INCLUDE") for ln in sim.circuit)
    assert nobs == 9 and ndet == (108 if rounds > 0 else 0) + (216 if rounds > 1 else 0) + 108


@pytest.mark.parametrize("rounds", [0, 1, 3])
def test_same_records_as_the_reference_generated_circuit(code225, rounds):
    """Under depolarizing_noise the noise sits at the measurement steps only, so the layer order of the
    CX / CZ gates cannot matter: with the same Philox stream this module's circuit and the circuit the
    reference generated must give bit-identical measurement records."""
    from oracle import pyoracle as po
    mine = compile_circuit(storage.build_storage_simulation(rounds, storage.depolarizing_noise(0.05, 0.05), code225).circuit)
    ref = compile_circuit(_golden(f"T{rounds}_depol"))
    assert (mine.num_measurements, mine.num_qubits, mine.num_sites) == (ref.num_measurements, ref.num_qubits, ref.num_sites)
    assert np.array_equal(mine.thresholds, ref.thresholds)
    a = po.circuit_sample(mine.ops, mine.thresholds, mine.num_qubits, mine.num_measurements, 9, 0, 300)
    b = po.circuit_sample(ref.ops, ref.thresholds, ref.num_qubits, ref.num_measurements, 9, 0, 300)
    assert np.array_equal(a, b)


def test_circuit_noise_structure_matches_the_reference_circuit(code225):
    """circuit_noise depends on the layering, so only structure and statistics can agree."""
    from oracle import pyoracle as po
    mine = compile_circuit(storage.build_storage_simulation(1, storage.circuit_noise(0.01, 0.02), code225).circuit)
    ref = compile_circuit(_golden("T1_circuit"))
    assert (mine.num_measurements, mine.num_qubits, mine.num_sites, mine.ops.shape) == \
        (ref.num_measurements, ref.num_qubits, ref.num_sites, ref.ops.shape)
    assert np.array_equal(np.bincount(mine.ops[:, 0] & 0xff), np.bincount(ref.ops[:, 0] & 0xff))
    a = po.circuit_sample(mine.ops, mine.thresholds, mine.num_qubits, mine.num_measurements, 1, 0, 4000)
    b = po.circuit_sample(ref.ops, ref.thresholds, ref.num_qubits, ref.num_measurements, 2, 0, 4000)
    for lo, hi in ((0, 108), (108, 216), (216, 441)):
        pa, pb = a[:, lo:hi].mean(), b[:, lo:hi].mean()
        assert abs(pa - pb) < 5 * np.sqrt(pa * (1 - pa) / a[:, lo:hi].size * 2 + 1e-12)


def test_code_capacity_rate():
    q = storage.depolarizing_noise(0.03, 0.01).code_capacity_flip_probability()
    assert abs(q - (0.02 + 0.01 - 2 * 0.02 * 0.01)) < 1e-15
    assert storage.trivial_noise().code_capacity_flip_probability() == 0.0

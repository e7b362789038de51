"""Random biregular Tanner graphs, girth bound and the HGP built on them (host builders; CPU only).

Mirrors the reference's ``tests/test_random_biregular_graph.py`` (degree smoke tests over its seeds and
shapes ``:21-42``, short-cycle removal ``:44-56``, the cycle search on a hexagon ``:58-63``) and the CSS
invariants of ``tests/test_hypergraph_product_code.py:5-34`` on this repo's numpy-only generators.
"""
import numpy as np
import pytest

from exp_ldpc_b200 import code_builders as cb

SEEDS = [0x59824c5a, 0x9dca707a, 0xe0218aa8, 0x81da8035, 0x63b16deb, 0x7dc89245, 0x1ab46afa, 0x5cc6d93e]
CASES = [(36, 3, 4), (12, 5, 6), (24, 7, 8), (30, 9, 10)]          # (data vertices, data degree, check degree)


def _biregular(a, data_degree, check_degree):
    assert set(np.unique(a).tolist()) <= {0, 1}
    assert (a.sum(axis=1) == data_degree).all() and (a.sum(axis=0) == check_degree).all()


@pytest.mark.parametrize("num_data,data_degree,check_degree", CASES)
@pytest.mark.parametrize("seed", SEEDS)
def test_random_biregular_matrix(num_data, data_degree, check_degree, seed):
    a = cb.random_biregular_matrix(num_data, data_degree, check_degree, seed=seed)
    assert a.shape == (num_data, num_data * data_degree // check_degree)
    _biregular(a, data_degree, check_degree)
    assert np.array_equal(a, cb.random_biregular_matrix(num_data, data_degree, check_degree, seed=seed))


def test_incompatible_degrees_raise():
    with pytest.raises(RuntimeError):
        cb.random_biregular_matrix(10, 3, 4)


@pytest.mark.parametrize("seed", SEEDS)
def test_remove_short_cycles(seed):
    a = cb.random_biregular_matrix(136, 3, 4, seed=seed)
    cb.remove_short_cycles(a, 4, seed=seed - 42, patience=10000)
    _biregular(a, 3, 4)
    assert cb.girth_exceeds(a, 4)
    # no two data vertices share two checks  <=>  no 4-cycle
    overlap = a.astype(np.int64) @ a.T.astype(np.int64)
    np.fill_diagonal(overlap, 0)
    assert overlap.max() <= 1


def test_cycle_search_on_a_hexagon():
    ring = np.zeros((3, 3), dtype=np.uint8)
    for i in range(3):
        ring[i, i] = ring[i, (i + 1) % 3] = 1
    assert cb.find_short_cycle(ring, 0, 2) is None
    length, edge = cb.find_short_cycle(ring, 0, 3)
    assert length == 6 and ring[edge]
    square = np.ones((2, 2), dtype=np.uint8)
    assert cb.find_short_cycle(square, 0, 2)[0] == 4 and not cb.girth_exceeds(square, 4)


def test_patience_is_enforced():
    with pytest.raises(RuntimeError):
        cb.remove_short_cycles(np.ones((3, 3), dtype=np.uint8), 4, seed=0, patience=50)     # K_{3,3}: nothing to swap with


@pytest.mark.parametrize("num_data,girth_bound", [(12, None), (24, 4)])
def test_biregular_hgp(num_data, girth_bound):
    nc = num_data * 3 // 4
    code = cb.biregular_hgp(num_data, 3, 4, seed=7, girth_bound=girth_bound)
    hx, hz = code.checks.x, code.checks.z
    n = num_data ** 2 + nc ** 2
    assert code.num_qubits == n and hx.shape == hz.shape == (num_data * nc, n)
    assert set(np.diff(hz.indptr).tolist()) == {7} and set(np.diff(hz.tocsc().indptr).tolist()) == {3, 4}
    assert not ((hx @ hz.T).toarray() % 2).any()
    lx, lz = code.logicals.x.astype(np.int64), code.logicals.z.astype(np.int64)
    k = code.num_logicals
    assert k >= (num_data - nc) ** 2 and lx.shape == lz.shape == (k, n)
    assert not ((hx @ lz.T) % 2).any() and not ((hz @ lx.T) % 2).any()
    assert np.array_equal((lz @ lx.T) % 2, np.eye(k, dtype=np.int64))


def test_random_test_hgp_shape():
    code = cb.random_test_hgp(compute_logicals=False)
    assert code.num_qubits == 36 ** 2 + 27 ** 2 == 2025
    assert code.checks.z.shape == (972, 2025) and set(np.diff(code.checks.z.indptr).tolist()) == {7}


def test_package_exports_the_reference_names():
    """python/qldpc/__init__.py:1-13 — what a user of the reference imports from the package root."""
    import exp_ldpc_b200 as q
    for name in ("random_test_hgp", "QuantumCodeChecks", "make_check_matrix", "build_storage_simulation", "biregular_hgp",
                 "remove_short_cycles", "qc_lifted_product_code", "lifted_product_code_pgl2", "lifted_product_code_cyclic",
                 "read_quantum_code", "write_quantum_code", "QuantumCode", "QuantumCodeLogicals", "CircuitTargets", "StorageSim",
                 "SpacetimeCode", "SpacetimeCodeSingleShot", "noise_model"):
        assert getattr(q, name) is not None, name
    assert callable(q.noise_model.trivial_noise) and callable(q.noise_model.depolarizing_noise)
    assert set(q.__all__) <= set(dir(q)) | set(q.__all__)
    with pytest.raises(AttributeError):
        q.no_such_name


def test_generate_hgp_code_script(tmp_path):
    """scripts/generate_hgp_code.py 4 3 12 (BASELINE configs[0]): qecc file + noiseless extraction circuit."""
    import os
    import subprocess
    import sys
    from exp_ldpc_b200.codes import load_code
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code_path, circ_path = tmp_path / "c.qecc", tmp_path / "c.stim"
    cmd = [sys.executable, os.path.join(root, "scripts", "generate_hgp_code.py"), "4", "3", "12", "--seed", "3",
           "--save_code", str(code_path), "--save_circuit", str(circ_path)]
    assert subprocess.run(cmd, capture_output=True, text=True, timeout=300).returncode == 0
    code = load_code(code_path)                                   # read_quantum_code validates the CSS conditions
    assert code.num_qubits == 225 and code.checks.z.shape == (108, 225) and code.num_logicals >= 9
    text = circ_path.read_text().splitlines()
    assert text[0].startswith("RZ 0 1 2") and sum(ln.startswith("MRX") for ln in text) == 2
    assert sum(ln.startswith("OBSERVABLE_INCLUDE") for ln in text) == code.num_logicals
    # an existing destination is refused, like the reference's script
    again = subprocess.run(cmd, capture_output=True, text=True, timeout=300)
    assert again.returncode != 0 and "already exists" in again.stdout

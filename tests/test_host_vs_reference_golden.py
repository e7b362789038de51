"""Host mirror (exp_ldpc_b200.codes / spacetime) against fixtures produced by RUNNING the reference
(tests/golden/make_golden.py): qecc IO, spacetime matrices, syndrome differencing, correction fold,
measurement-record layout.  CPU only."""
import io
import json
import os

import numpy as np
import pytest
from scipy import sparse

from conftest import CODES, GOLDEN
from exp_ldpc_b200.codes import (QuantumCode, QuantumCodeChecks, QuantumCodeLogicals, load_code, make_check_matrix,
                                 read_quantum_code, write_quantum_code)
from exp_ldpc_b200.spacetime import MeasurementLayout, SpacetimeCode, SpacetimeCodeSingleShot


def _csr(ptr, idx, shape):
    return sparse.csr_matrix((np.ones(len(idx), dtype=np.uint32), idx, ptr), shape=tuple(shape))


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(GOLDEN, "spacetime_n225.npz"))


def test_reader_matches_reference_writer(code225):
    # file emitted by the reference's write_quantum_code, read by our reader
    with open(os.path.join(GOLDEN, "hgp_34_n225_refwriter.qecc")) as fh:
        ref = read_quantum_code(fh, validate_stabilizer_code=True)
    assert (ref.checks.x != code225.checks.x).nnz == 0
    assert (ref.checks.z != code225.checks.z).nnz == 0
    assert np.array_equal(ref.logicals.x, code225.logicals.x)
    assert np.array_equal(ref.logicals.z, code225.logicals.z)


def test_writer_matches_reference_writer(code225):
    buf = io.StringIO()
    write_quantum_code(buf, code225)
    with open(os.path.join(GOLDEN, "hgp_34_n225_refwriter.qecc")) as fh:
        assert buf.getvalue() == fh.read()


def test_io_roundtrip(code225):   # tests/test_quantum_code_io.py:6-22
    buf = io.StringIO()
    write_quantum_code(buf, code225)
    buf.seek(0)
    back = read_quantum_code(buf, validate_stabilizer_code=True)
    assert (back.checks.x != code225.checks.x).nnz == 0 and (back.checks.z != code225.checks.z).nnz == 0
    assert back.num_qubits == code225.num_qubits
    assert np.array_equal(back.logicals.x, code225.logicals.x) and np.array_equal(back.logicals.z, code225.logicals.z)


@pytest.mark.parametrize("text,msg", [
    ("qecc 4 1 1\n", "Invalid header"),
    ("qecc 2 2 2 0\n0 X\n1 X\n0 Z\n1 Z\n", "overconstrained"),
    ("qecc 4 1 1 0\n0 1 Q\n", "Invalid check/logical type"),
    ("qecc 4 1 1 0\n0 9 X\n0 1 Z\n", "Out of bounds"),
    ("qecc 4 1 1 0\n0 1 X\n", "Number of checks does not match"),
    ("qecc 4 1 1 1\n0 1 X\n0 1 Z\n0 LX\n", "Number of X and Z logicals"),
    ("qecc 4 1 1 2\n0 1 X\n0 1 Z\n0 1 LX\n0 1 LZ\n", "Parsed number of logicals"),
    ("qecc 4 1 1 0\n0 1 X\n1 2 Z\n", "abelian"),
])
def test_reader_errors(text, msg):   # quantum_code_io.py:16-60 raise RuntimeError
    with pytest.raises(RuntimeError, match=msg):
        read_quantum_code(io.StringIO(text))


def test_reader_skips_comments():
    code = read_quantum_code(io.StringIO("c hello\nqecc 4 1 1 0\nc mid\n0 1 X\n0 1 Z\n\n"))
    assert code.num_qubits == 4 and code.checks.x.shape == (1, 4)


def test_container_errors():
    with pytest.raises(ValueError):
        QuantumCodeChecks(sparse.csr_matrix(np.eye(3, dtype=np.uint32)), sparse.csr_matrix(np.eye(4, dtype=np.uint32)))
    with pytest.raises(TypeError):
        QuantumCodeChecks(sparse.csr_matrix(np.eye(3)), sparse.csr_matrix(np.eye(3)))
    with pytest.raises(ValueError):
        QuantumCodeLogicals(np.zeros((1, 3), dtype=np.uint32), np.zeros((2, 3), dtype=np.uint32))
    c = QuantumCode(QuantumCodeChecks(make_check_matrix([[0, 1]], 3), make_check_matrix([[0, 1]], 3)))
    assert c.num_logicals == 0 and c.logicals.z.shape == (0, 3)
    assert not c.checks.z.data.flags.writeable


def test_checks_are_canonical_for_any_index_dtype(code225):
    hz = code225.checks.z
    for idx_t in (np.int32, np.int64):
        m = sparse.csr_array((hz.data, hz.indices.astype(idx_t), hz.indptr.astype(idx_t)), shape=hz.shape)
        c = QuantumCodeChecks(m, m)
        assert c.z.indices.dtype == np.int32 and (c.z != hz).nnz == 0
    coo = hz.tocoo()
    perm = np.random.default_rng(0).permutation(coo.nnz)
    shuffled = sparse.coo_matrix((coo.data[perm], (coo.row[perm], coo.col[perm])), shape=hz.shape)
    assert (QuantumCodeChecks(shuffled, shuffled).z != hz).nnz == 0


@pytest.mark.parametrize("T", [0, 1, 3])
def test_spacetime_matrix_matches_reference(gold, T):
    hz = _csr(gold["hz_indptr"], gold["hz_indices"], gold["hz_shape"])
    st = SpacetimeCode(hz, T)
    ref = _csr(gold[f"st{T}_indptr"], gold[f"st{T}_indices"], gold[f"st{T}_shape"])
    m, n = hz.shape
    assert st.spacetime_check_matrix.shape == ((T + 1) * m, (T + 1) * n + T * m)
    assert (st.spacetime_check_matrix != ref).nnz == 0
    assert st.spacetime_check_matrix.nnz == (T + 1) * hz.nnz + 2 * T * m


@pytest.mark.parametrize("T", [0, 1, 3])
def test_spacetime_syndrome_and_fold_match_reference(gold, T):
    hz = _csr(gold["hz_indptr"], gold["hz_indices"], gold["hz_shape"])
    st = SpacetimeCode(hz, T)
    hist = gold[f"st{T}_hist"]
    synd = st.syndrome_from_history(lambda t: hist[t], gold[f"st{T}_readout"])
    assert synd.dtype == np.float64 and np.array_equal(synd, gold[f"st{T}_synd"])
    assert synd.shape[0] == st.spacetime_check_matrix.shape[0]      # tests/test_spacetime_code.py:14-24
    assert np.array_equal(st.final_correction(gold[f"st{T}_fault"]), gold[f"st{T}_fold"])
    prior = np.zeros(st.spacetime_check_matrix.shape[1])
    st.data_bits(prior)[:] = 0.25
    st.measurement_bits(prior)[:] = 0.5
    assert np.array_equal(prior, gold[f"st{T}_prior_quirk"])          # the T*m split quirk, kept


def test_single_shot_matrix_matches_reference(gold):
    hz = _csr(gold["hz_indptr"], gold["hz_indices"], gold["hz_shape"])
    ss = SpacetimeCodeSingleShot(hz)
    ref = _csr(gold["ss_indptr"], gold["ss_indices"], gold["ss_shape"])
    assert (ss.spacetime_check_matrix != ref).nnz == 0
    x = np.arange(hz.shape[1] + hz.shape[0])
    assert np.array_equal(ss.data_bits(x), x[:hz.shape[1]]) and np.array_equal(ss.measurement_bits(x), x[hz.shape[1]:])


@pytest.mark.parametrize("T", [0, 1, 3])
def test_measurement_layout_matches_reference(code225, T):
    with open(os.path.join(GOLDEN, "reference_meta.json")) as fh:
        ref = json.load(fh)["storage_layout"][str(T)]
    lay = MeasurementLayout(code225.num_qubits, code225.checks.x.shape[0], code225.checks.z.shape[0], T)
    assert lay.record_length == ref["record_length"]
    rec = np.arange(lay.record_length)
    d = lay.data_view(rec)
    assert [int(d[0]), int(d[-1])] == ref["data_view"] and d.shape[0] == code225.num_qubits
    for t in range(T):
        z = lay.measurement_view(t, False, rec)
        x = lay.measurement_view(t, True, rec)
        assert [int(z[0]), int(z[-1])] == ref["z_views"][t] and [int(x[0]), int(x[-1])] == ref["x_views"][t]
    # views, not copies (tests/test_storage_sim.py:119-131)
    buf = np.zeros(lay.record_length, dtype=np.int64)
    lay.data_view(buf)[:] = 1
    assert buf.sum() == code225.num_qubits


def test_pinned_codes_have_expected_parameters():
    for name, n, m, k in [("hgp_34_n225", 225, 108, 9), ("hgp_34_n1600", 1600, 768, 64),
                          ("hgp_34_n6400", 6400, 3072, 256), ("hgp_34_n2025_test", 2025, 972, 81)]:
        c = load_code(os.path.join(CODES, name + ".qecc"))   # validates HxHz^T = 0 and logical commutation
        assert (c.num_qubits, c.checks.z.shape[0], c.checks.x.shape[0], c.num_logicals) == (n, m, m, k)
        # tests/test_hypergraph_product_code.py:5-34
        assert np.array_equal((c.logicals.z.astype(np.int64) @ c.logicals.x.T.astype(np.int64)) % 2, np.eye(k, dtype=np.int64))

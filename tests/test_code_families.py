"""Known answers of the non-HGP code families of BASELINE.json configs 3-4 (host builders; CPU).

Mirrors the reference's own known-answer tests:
  tests/test_qc_lifted_product_code.py:4-9        PK'20 shifts, l = 31      -> [[1054, 140]]
  tests/test_matrix_lifted_product_code.py:16-41  PK'19 B3 over Z_127       -> [[1270, 28]]
  tests/test_matrix_lifted_product_code.py:43-62  A5 = PSL(2,5) lift        -> n = 2700, k = 560
and adds the [[882, 24]] code BASELINE.json quotes for config 3 (PK'19 B1 over Z_63).
"""
import numpy as np
import pytest

from exp_ldpc_b200 import code_builders as cb


@pytest.mark.parametrize("name,n,k,row_w,col_w", [
    ("pk20_qc_1054", 1054, 140, {8}, {3, 5}),
    ("pk19_b3_1270", 1270, 28, {6}, {3}),
    ("pk19_b1_882", 882, 24, {6}, {3}),
    ("a5_lift_2700", 2700, 560, {9}, {3, 6}),
])
def test_known_parameters(name, n, k, row_w, col_w):
    code = getattr(cb, name)()
    hx, hz = code.checks.x, code.checks.z
    assert code.num_qubits == n and code.num_logicals == k
    assert set(np.diff(hz.indptr).tolist()) == row_w
    assert set(np.diff(hz.tocsc().indptr).tolist()) == col_w
    # CSS conditions of tests/test_hypergraph_product_code.py:5-34
    assert not ((hx @ hz.T).toarray() % 2).any()
    lx, lz = code.logicals.x.astype(np.int64), code.logicals.z.astype(np.int64)
    assert not ((hx @ lz.T) % 2).any() and not ((hz @ lx.T) % 2).any()
    assert np.array_equal((lz @ lx.T) % 2, np.eye(k, dtype=np.int64))

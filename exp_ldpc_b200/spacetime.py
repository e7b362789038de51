"""Decoding-problem builders: the matrices the decoder sees and the per-shot pre/post steps.

Host-side mirror of ``python/qldpc/spacetime_code.py`` (names and semantics kept):

* :class:`SpacetimeCode` — ``H_st = [ blockdiag(H x (T+1)) | M ]`` (``spacetime_code.py:46-75``);
* :class:`SpacetimeCodeSingleShot` — ``[H | I_m]`` (``spacetime_code.py:10-37``);
* :func:`spacetime_syndrome` — history + transversal readout -> differenced syndrome
  (``spacetime_code.py:98-119``);
* ``final_correction`` — XOR-fold of the ``T+1`` data blocks (``spacetime_code.py:81-84``).

These are the *single-shot, host* versions used for set-up, for the golden-fixture tests and for
the ``readout_correction`` drop-ins.  The batched device versions of the syndrome differencing
and the fold live in ``csrc/`` and are reached through :mod:`exp_ldpc_b200.decoders`.

Deviation kept on purpose: the reference's ``data_bits``/``measurement_bits`` views of
``SpacetimeCode`` split at ``T*m`` (``spacetime_code.py:75,86-92``) instead of at ``(T+1)*n``.
``data_bits``/``measurement_bits`` here reproduce that split bit for bit (so priors written
through them land where the reference puts them); :meth:`SpacetimeCode.true_data_bits` gives the
geometrically correct split.
"""
from __future__ import annotations

from typing import Callable

import numpy as np
from scipy import sparse

from .codes import canonical_csr

__all__ = ["SpacetimeCode", "SpacetimeCodeSingleShot", "spacetime_syndrome", "MeasurementLayout"]


class SpacetimeCodeSingleShot:
    """``[H | I]``: one extra fault bit per check for a measurement flip."""

    def __init__(self, check_matrix):
        h = canonical_csr(check_matrix)
        m, n = h.shape
        self._datablock_size = n
        self.spacetime_check_matrix = sparse.hstack(
            [h, sparse.identity(m, dtype=np.uint32, format="csr")], format="csr").astype(np.uint32)

    def final_correction(self, x):
        return self.data_bits(x)

    def data_bits(self, x):
        return x[:self._datablock_size]

    def measurement_bits(self, x):
        return x[self._datablock_size:]


class SpacetimeCode:
    """Multi-round code tracking syndrome differences.

    Column order: data faults of round 0 .. round T (``n`` each), then ``T*m`` measurement
    faults; measurement fault ``t*m + j`` touches rows ``t*m + j`` and ``(t+1)*m + j``.
    """

    def __init__(self, check_matrix, num_rounds: int):
        h = canonical_csr(check_matrix)
        m, n = h.shape
        T = int(num_rounds)
        if T < 0:
            raise ValueError("num_rounds must be >= 0")
        blocks = sparse.block_diag([h] * (T + 1), format="csr") if T > 0 else h
        t_idx, j_idx = np.divmod(np.arange(T * m, dtype=np.int64), m) if T > 0 else (np.zeros(0, np.int64),) * 2
        col = np.arange(T * m, dtype=np.int64)
        rows = np.concatenate([t_idx * m + j_idx, (t_idx + 1) * m + j_idx])
        cols = np.concatenate([col, col])
        meas = sparse.coo_matrix(
            (np.ones(rows.shape[0], dtype=np.uint32), (rows, cols)), shape=((T + 1) * m, T * m), dtype=np.uint32)
        self.spacetime_check_matrix = sparse.hstack([blocks, meas], format="csr").astype(np.uint32)
        self._check_matrix = h
        self._num_rounds = T
        self._datablock_size = T * m  # sic — reference quirk, see module docstring

    @property
    def num_rounds(self) -> int:
        return self._num_rounds

    def syndrome_from_history(self, history: Callable[[int], np.ndarray], readout) -> np.ndarray:
        return spacetime_syndrome(self._num_rounds, self._check_matrix, history, readout)

    def final_correction(self, spacetime_correction):
        n = self._check_matrix.shape[1]
        x = np.asarray(spacetime_correction)
        acc = np.zeros(n, dtype=np.int64)
        for t in range(self._num_rounds + 1):
            acc += x[t * n:(t + 1) * n].astype(np.int64)
        return acc % 2

    def data_bits(self, x):
        return x[:self._datablock_size]

    def measurement_bits(self, x):
        return x[self._datablock_size:]

    def true_data_bits(self, x):
        n = self._check_matrix.shape[1]
        return x[:(self._num_rounds + 1) * n]

    def true_measurement_bits(self, x):
        n = self._check_matrix.shape[1]
        return x[(self._num_rounds + 1) * n:]


def spacetime_syndrome(rounds: int, check_matrix, syndrome_history: Callable[[int], np.ndarray], readout) -> np.ndarray:
    """Rows ``t < T``: measured syndrome of round t; row T: ``H @ readout mod 2``; then every row
    ``t >= 1`` is XORed with row ``t-1`` of the *raw* table.  float64 output like the reference."""
    h = check_matrix
    m = h.shape[0]
    table = np.zeros((rounds + 1, m), dtype=np.float64)
    for t in range(rounds):
        table[t] = np.asarray(syndrome_history(t))
    table[rounds] = (h @ np.asarray(readout)) % 2
    out = table.copy()
    out[1:] = (table[1:] + table[:-1]) % 2
    return out.reshape(-1)


class MeasurementLayout:
    """Per-shot measurement-record layout of the storage simulation
    (``storage_sim.py:187-196``): ``[round 0: X(mx) | Z(mz)] ... [round T-1] [data(n)]``.

    ``measurement_view`` / ``data_view`` return *views* (the reference's tests require that,
    ``tests/test_storage_sim.py:119-131``).
    """

    def __init__(self, num_data: int, x_check_count: int, z_check_count: int, rounds: int):
        self.num_data = int(num_data)
        self.x_check_count = int(x_check_count)
        self.z_check_count = int(z_check_count)
        self.rounds = int(rounds)

    @property
    def record_length(self) -> int:
        return self.rounds * (self.x_check_count + self.z_check_count) + self.num_data

    def measurement_view(self, round_index: int, get_x_checks: bool, measurement_vector):
        off = (self.x_check_count + self.z_check_count) * round_index
        if get_x_checks:
            return measurement_vector[off:off + self.x_check_count]
        off += self.x_check_count
        return measurement_vector[off:off + self.z_check_count]

    def data_view(self, measurement_vector):
        off = (self.x_check_count + self.z_check_count) * self.rounds
        return measurement_vector[off:off + self.num_data]

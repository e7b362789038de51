"""Code containers and the ``qecc`` text format — the input types of the decode boundary.

Host-side mirror of the reference's ``python/qldpc/qecc_util.py:19-118`` (containers) and
``python/qldpc/quantum_code_io.py:6-71`` (file format), same names, argument meaning and error
behaviour, so that objects produced here can be handed to the decoders in
:mod:`exp_ldpc_b200.decoders` exactly where the reference hands its own to ``ldpc``.

File format (``quantum_code_io.py:15-35,64-71``)::

    c comment lines start with 'c'
    qecc <n qubits> <n X checks> <n Z checks> <n logicals>
    <qubit> <qubit> ... X|Z|LX|LZ          one row per generator / logical

Checks are canonical CSR ``uint32`` (sorted, deduplicated column indices — this *defines the edge
order* every decoder sees, ``qecc_util.py:29-43``); logicals are dense ``uint32`` ``k x n``.
"""
from __future__ import annotations

import warnings
from dataclasses import dataclass
from typing import Iterable, Sequence

import numpy as np
from scipy import sparse

__all__ = [
    "QuantumCodeChecks", "QuantumCodeLogicals", "QuantumCode", "make_check_matrix",
    "read_quantum_code", "write_quantum_code", "load_code", "canonical_csr",
]


def _require_integral(m, what: str) -> None:
    # qecc_util.py:12-17: non-integer dtype is a TypeError, signed integers only warn
    if not np.issubdtype(m.dtype, np.integer):
        raise TypeError(f"{what}: got numpy object with non-integral dtype {m.dtype}")
    if np.issubdtype(m.dtype, np.signedinteger):
        warnings.warn(f"{what}: signed integer dtype; arithmetic mod 2 may overflow", stacklevel=3)


def canonical_csr(h) -> sparse.csr_matrix:
    """Any scipy sparse / dense 0-1 matrix -> CSR with sorted unique indices, int32 index arrays.

    Accepts ``csr_matrix``/``csr_array``/COO and int32 or int64 indices (scipy >= 1.14 yields
    ``csr_array`` with int64; the reference's pinned 1.13 yields int32 ``csr_matrix``).
    Entries are reduced mod 2 and explicit zeros dropped, so the result is a pure support pattern.
    """
    if not sparse.issparse(h):
        h = sparse.csr_matrix(np.asarray(h))
    h = sparse.csr_matrix(h)  # copies; also converts sparray -> spmatrix
    h.sum_duplicates()
    h.sort_indices()
    h.data = (h.data.astype(np.int64) & 1).astype(np.uint32)
    h.eliminate_zeros()
    h.indices = h.indices.astype(np.int32, copy=False)
    h.indptr = h.indptr.astype(np.int32, copy=False)
    return h


class QuantumCodeChecks:
    """X and Z parity checks (``qecc_util.py:19-51``)."""

    __slots__ = ("x", "z")

    def __init__(self, x, z):
        for name, h in (("x", x), ("z", z)):
            if not sparse.issparse(h):
                h = sparse.csr_matrix(np.asarray(h))
            _require_integral(h, f"{name} checks")
            h = canonical_csr(h)
            h.data.flags.writeable = False
            object.__setattr__(self, name, h)
        if self.x.shape[1] != self.z.shape[1]:
            raise ValueError("x and z checks act on an inconsistent number of qubits")

    def __setattr__(self, key, value):  # frozen, like the reference dataclass
        raise AttributeError("QuantumCodeChecks is immutable")

    @property
    def num_qubits(self) -> int:
        return self.x.shape[1]


class QuantumCodeLogicals:
    """Dense logical operators, one row each (``qecc_util.py:53-91``)."""

    __slots__ = ("x", "z")

    def __init__(self, x, z):
        x = np.array(x, copy=True)
        z = np.array(z, copy=True)
        if x.ndim != 2 or z.ndim != 2:
            raise ValueError("logicals must be 2-D (k x n)")
        _require_integral(x, "x logicals")
        _require_integral(z, "z logicals")
        if x.shape[1] != z.shape[1]:
            raise ValueError("x and z logicals act on an inconsistent number of qubits")
        if x.shape[0] != z.shape[0]:
            raise ValueError("Number of provided X and Z logical operators mismatch")
        x.flags.writeable = False
        z.flags.writeable = False
        object.__setattr__(self, "x", x)
        object.__setattr__(self, "z", z)

    def __setattr__(self, key, value):
        raise AttributeError("QuantumCodeLogicals is immutable")

    @property
    def num_qubits(self) -> int:
        return self.x.shape[1]

    @property
    def num_logicals(self) -> int:
        return self.x.shape[0]

    @classmethod
    def empty(cls, num_qubits: int) -> "QuantumCodeLogicals":
        # the reference's version is a broken classmethod (qecc_util.py:86-91); this one works
        z = np.zeros((0, num_qubits), dtype=np.uint32)
        return cls(z, z.copy())


class QuantumCode:
    """Checks + logicals (``qecc_util.py:94-118``)."""

    __slots__ = ("checks", "logicals")

    def __init__(self, checks: QuantumCodeChecks, logicals: QuantumCodeLogicals | None = None):
        if logicals is None:
            logicals = QuantumCodeLogicals.empty(checks.num_qubits)
        if checks.num_qubits != logicals.num_qubits:
            raise ValueError("Number of qubits for checks and logicals is inconsistent")
        object.__setattr__(self, "checks", checks)
        object.__setattr__(self, "logicals", logicals)

    def __setattr__(self, key, value):
        raise AttributeError("QuantumCode is immutable")

    @property
    def num_qubits(self) -> int:
        return self.checks.num_qubits

    @property
    def num_logicals(self) -> int:
        return self.logicals.num_logicals


def make_check_matrix(rows: Sequence[Iterable[int]], num_qubits: int) -> sparse.csr_matrix:
    """Rows given as supports -> CSR uint32 (``qecc_util.py:146-149``)."""
    rows = [list(r) for r in rows]
    lens = np.fromiter((len(r) for r in rows), dtype=np.int64, count=len(rows))
    indptr = np.zeros(len(rows) + 1, dtype=np.int64)
    np.cumsum(lens, out=indptr[1:])
    indices = np.fromiter((v for r in rows for v in r), dtype=np.int64, count=int(indptr[-1]))
    data = np.ones(indices.shape[0], dtype=np.uint32)
    m = sparse.csr_matrix((data, indices, indptr), shape=(len(rows), num_qubits), dtype=np.uint32)
    return m


_ROW_KINDS = ("X", "Z", "LX", "LZ")


def read_quantum_code(stream, validate_stabilizer_code=None) -> QuantumCode:
    """Parse a ``qecc`` file (``quantum_code_io.py:6-62``); every malformed input raises
    ``RuntimeError`` exactly where the reference does."""
    if validate_stabilizer_code is None:
        validate_stabilizer_code = True
    toks = [ln.split() for ln in stream.readlines() if not ln.startswith("c")]
    toks = [t for t in toks if t]
    if not toks or toks[0][0] != "qecc" or len(toks[0]) != 5:
        raise RuntimeError("Invalid header. Expected qecc <# qubits> <# X checks> <# Z checks> <# logicals>")
    n, nx, nz, k = (int(v) for v in toks[0][1:])
    if nx + nz > n:
        raise RuntimeError(f"Code overconstrained. Got {nx + nz} checks on {n} qubits")
    rows = {kind: [] for kind in _ROW_KINDS}
    for t in toks[1:]:
        kind = t[-1]
        if kind not in rows:
            raise RuntimeError(f"Invalid check/logical type in line: \n {t}")
        support = [int(v) for v in t[:-1]]
        if any(v >= n for v in support):
            raise RuntimeError(f"Out of bounds check support: \n {t}")
        rows[kind].append(support)
    if len(rows["X"]) + len(rows["Z"]) != nx + nz:
        raise RuntimeError(
            f"Number of checks does not match header. Expected {nx} + {nz}. "
            f"Got {len(rows['X'])} + {len(rows['Z'])}")
    if len(rows["LZ"]) != len(rows["LX"]):
        raise RuntimeError(
            f"Number of X and Z logicals does not match: {len(rows['LX'])} X logicals and "
            f"{len(rows['LZ'])} Z logicals")
    if len(rows["LZ"]) != k:
        raise RuntimeError(
            f"Parsed number of logicals does not match header. Expected {k}. Got {len(rows['LZ'])}")
    checks = QuantumCodeChecks(make_check_matrix(rows["X"], n), make_check_matrix(rows["Z"], n))
    logicals = QuantumCodeLogicals(
        make_check_matrix(rows["LX"], n).toarray(), make_check_matrix(rows["LZ"], n).toarray())
    if validate_stabilizer_code is True:
        if ((checks.x @ checks.z.T).data % 2).any():
            raise RuntimeError("X and Z checks do not generate an abelian group")
        if logicals.num_logicals > 0:
            if ((checks.x @ logicals.z.T) % 2).any():
                raise RuntimeError("Z logicals do not commute with X checks")
            if ((checks.z @ logicals.x.T) % 2).any():
                raise RuntimeError("X logicals do not commute with Z checks")
    return QuantumCode(checks, logicals)


def write_quantum_code(stream, code: QuantumCode) -> None:
    """Emit the ``qecc`` text (``quantum_code_io.py:64-71``): header, X rows, Z rows, LZ, LX."""
    cx, cz = code.checks.x, code.checks.z
    stream.write(f"qecc {code.num_qubits} {cx.shape[0]} {cz.shape[0]} {code.num_logicals}\n")
    for kind, h in (("X", cx), ("Z", cz)):
        ptr, idx = h.indptr, h.indices
        for r in range(h.shape[0]):
            cols = " ".join(str(int(c)) for c in idx[ptr[r]:ptr[r + 1]])
            stream.write(f"{cols} {kind}\n")
    for kind, dense in (("LZ", code.logicals.z), ("LX", code.logicals.x)):
        for r in range(dense.shape[0]):
            cols = " ".join(str(int(c)) for c in np.flatnonzero(dense[r]))
            stream.write(f"{cols} {kind}\n")


def load_code(path) -> QuantumCode:
    """``load_code`` of ``misc/_experiment.py:231-235``: a path, or — as there — parsed command-line
    arguments whose ``code`` attribute is the path."""
    if not isinstance(path, (str, bytes)) and not hasattr(path, "__fspath__") and hasattr(path, "code"):
        path = path.code
    with open(path, "r") as fh:
        return read_quantum_code(fh, validate_stabilizer_code=True)

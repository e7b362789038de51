"""Bit-packed GF(2) linear algebra used at set-up time (never on the per-shot path).

The reference delegates this to the third-party ``galois`` package
(``python/qldpc/homological_product_code.py:6-60``, ``python/qldpc/linalg.py:93-99``); that package
is not available here, so rank / null space / logical-operator extraction are restated on rows
packed into ``uint64`` words.  All functions take and return dense ``uint8`` 0/1 matrices at the
boundary so callers never see the packing.
"""
from __future__ import annotations

import numpy as np

__all__ = [
    "pack_rows", "unpack_rows", "row_reduce", "rank", "null_space", "row_space_complement",
    "css_logicals", "in_row_space",
]


def _as_bits(a) -> np.ndarray:
    if hasattr(a, "toarray"):
        a = a.toarray()
    a = np.asarray(a)
    if a.ndim != 2:
        raise ValueError("expected a 2-D matrix")
    return (a & 1).astype(np.uint8) if a.dtype.kind in "iu" else (np.asarray(a, dtype=np.int64) & 1).astype(np.uint8)


def pack_rows(a) -> np.ndarray:
    """0/1 matrix ``r x c`` -> ``r x ceil(c/64)`` uint64, bit ``j%64`` of word ``j//64`` = column j."""
    a = _as_bits(a)
    r, c = a.shape
    w = (c + 63) // 64
    padded = np.zeros((r, w * 64), dtype=np.uint8)
    padded[:, :c] = a
    by = np.ascontiguousarray(np.packbits(padded, axis=1, bitorder="little"))
    return by.view("<u8").astype(np.uint64).reshape(r, w)


def unpack_rows(p: np.ndarray, ncols: int) -> np.ndarray:
    r, w = p.shape
    shifts = np.arange(64, dtype=np.uint64)
    bits = ((p[:, :, None] >> shifts) & np.uint64(1)).astype(np.uint8)
    return bits.reshape(r, w * 64)[:, :ncols]


def row_reduce(a, ncols: int | None = None):
    """Reduced row-echelon form over GF(2).

    Returns ``(R, pivots)`` with ``R`` the RREF restricted to its non-zero rows (dense uint8) and
    ``pivots`` the pivot column of each row.  ``ncols`` limits pivot search to the first columns
    (the augmented-matrix idiom of ``homological_product_code.py:31-33``).
    """
    a = _as_bits(a)
    nrows, nc = a.shape
    limit = nc if ncols is None else ncols
    p = pack_rows(a)
    pivots = []
    r = 0
    one = np.uint64(1)
    for j in range(limit):
        if r == nrows:
            break
        w, b = j >> 6, np.uint64(j & 63)
        col = (p[:, w] >> b) & one
        cand = np.nonzero(col[r:])[0]
        if cand.size == 0:
            continue
        i = r + int(cand[0])
        if i != r:
            p[[r, i]] = p[[i, r]]
            col[[r, i]] = col[[i, r]]
        col[r] = 0
        rows = np.nonzero(col)[0]
        if rows.size:
            p[rows] ^= p[r]
        pivots.append(j)
        r += 1
    if ncols is None:
        out = unpack_rows(p[:r], nc)
    else:
        out = unpack_rows(p, nc)  # keep every row: caller reads the augmented part
    return out, np.asarray(pivots, dtype=np.int64)


def rank(a) -> int:
    return int(row_reduce(a)[1].size)


def null_space(a) -> np.ndarray:
    """Basis (rows) of ``{x : a x = 0}``; one vector per free column, RREF-derived."""
    a = _as_bits(a)
    nc = a.shape[1]
    R, piv = row_reduce(a)
    free = np.setdiff1d(np.arange(nc), piv)
    basis = np.zeros((free.size, nc), dtype=np.uint8)
    for t, f in enumerate(free):
        basis[t, f] = 1
        if piv.size:
            basis[t, piv] = R[:, f]
    return basis


def in_row_space(basis, v) -> bool:
    """True iff ``v`` is a GF(2) combination of the rows of ``basis``."""
    basis = _as_bits(basis)
    v = _as_bits(np.atleast_2d(v))
    return rank(np.vstack([basis, v])) == rank(basis)


def row_space_complement(image, kernel) -> np.ndarray:
    """Rows of ``kernel`` that extend ``rowspace(image)`` to ``rowspace(image)+rowspace(kernel)``.

    Restates the basis-extension step of ``compute_homology_reps``
    (``homological_product_code.py:6-21``): stack ``[image; kernel]`` as *columns*, row-reduce,
    and keep the kernel vectors whose column carries a pivot.
    """
    image = _as_bits(image)
    kernel = _as_bits(kernel)
    if kernel.shape[0] == 0:
        return kernel
    aug = np.vstack([image, kernel]).T  # columns = candidate vectors
    _, piv = row_reduce(aug)
    keep = piv[piv >= image.shape[0]] - image.shape[0]
    return kernel[keep]


def css_logicals(hx, hz):
    """Logical operators of the CSS code (Hx, Hz): returns ``(Lx, Lz)`` uint32 ``k x n`` with
    ``Hz Lx^T = 0``, ``Hx Lz^T = 0`` and ``Lz Lx^T = I`` (the pairing that
    ``compute_logical_pairs``, ``homological_product_code.py:23-35`` establishes).
    """
    hx = _as_bits(hx)
    hz = _as_bits(hz)
    n = hx.shape[1]
    if hz.shape[1] != n:
        raise ValueError("x and z checks act on an inconsistent number of qubits")
    hx_rows = row_reduce(hx)[0]
    hz_rows = row_reduce(hz)[0]
    lx = row_space_complement(hx_rows, null_space(hz))
    lz = row_space_complement(hz_rows, null_space(hx))
    k = lx.shape[0]
    if lz.shape[0] != k:
        raise ValueError("X and Z logical counts differ; checks do not commute?")
    if k == 0:
        return lx.astype(np.uint32), lz.astype(np.uint32)
    # make (Lz Lx^T) the identity by row operations on Lz
    inner = (lz.astype(np.int64) @ lx.T.astype(np.int64)) & 1
    aug = np.hstack([inner.astype(np.uint8), lz])
    red, piv = row_reduce(aug, ncols=k)
    if piv.size != k:
        raise ValueError("logical pairing matrix is singular")
    lz = red[:k, k:]
    return lx.astype(np.uint32), lz.astype(np.uint32)

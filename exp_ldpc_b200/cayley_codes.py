"""Lifted product codes over Cayley graphs with local codes (BASELINE.json configs[3]; set-up time
only — nothing here is on the decode hot path).

Restates ``python/qldpc/lifted_product_code.py`` of the reference without ``galois`` / ``networkx``:

* finite fields GF(p^k) as addition / multiplication tables (:class:`Field`), the groups
  PGL(2,q), PSL(2,q) and Z_q^m as Cayley tables (:class:`FiniteGroup`) — the reference walks
  hashable matrix objects with a DFS (``lifted_product_code.py:19-137, 192-219``);
* the Morgenstern generators of PGL(2, q^i), q = 2^l (``:159-190``);
* :func:`lifted_product_code` (``:249-404``).  The reference enumerates the supports of every
  check and qubit as Python objects; here the same complex is written as what it is — the lifted
  product of two matrices over the group algebra,

      A[(v, r1), e] = h1[r1, in_idx_v(e)] * g_e  (v = head of e)  +  h1[r1, out_idx_v(e)] * 1  (v = tail of e)
      B[e, (v, r2)] = h2[r2, in_idx_v(e)] * g_e^-1 (v = head of e) +  h2[r2, out_idx_v(e)] * 1  (v = tail of e)

  with A acting from the left and B from the right, over the base graph D_w (two vertices, the w
  generators as parallel edges 0 -> 1: the "double cover") or B_w (one vertex, w loops) — and handed
  to :func:`code_builders.lifted_product`:  X checks = E x V, qubits = V x V + E x E, Z checks = V x E.

The reference iterates a Python ``set`` of group elements hashed by ``bytes`` (order changes with
PYTHONHASHSEED), so its qubit order is not reproducible even by itself; codes built here agree with
it up to that relabelling (same generators, same local codes for a given seed: both use
``numpy.random.default_rng``), and are pinned by their parameters (n, the CSS condition, k by rank)
as ``tests/test_lifted_product_code.py`` of the reference does.
"""
from __future__ import annotations

import itertools
import warnings
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import gf2
from .code_builders import lifted_product
from .codes import QuantumCode

__all__ = [
    "Field", "FiniteGroup", "pgl2_group", "psl2_group", "get_psl2", "zqm_group", "random_check_matrix",
    "random_abelian_generators", "morgenstern_matrices", "morgenstern_generators", "lifted_product_code",
    "lifted_product_code_cyclic", "lifted_product_code_pgl2", "psl2_tanner_lp_4320",
]


# ---------------------------------------------------------------------------------------------
# finite fields as tables
# ---------------------------------------------------------------------------------------------

# Conway polynomials (low-to-high coefficients), the defaults of the reference's field library,
# so that element labels (integers = base-p digits of the polynomial) mean the same thing here.
_CONWAY = {
    (2, 1): (1, 1), (2, 2): (1, 1, 1), (2, 3): (1, 1, 0, 1), (2, 4): (1, 1, 0, 0, 1), (2, 5): (1, 0, 1, 0, 0, 1),
    (2, 6): (1, 1, 0, 1, 1, 0, 1), (2, 7): (1, 1, 0, 0, 0, 0, 0, 1), (2, 8): (1, 0, 1, 1, 1, 0, 0, 0, 1),
    (3, 1): (1, 1), (3, 2): (2, 2, 1), (3, 3): (1, 2, 0, 1), (3, 4): (2, 0, 0, 2, 1),
    (5, 1): (3, 1), (5, 2): (2, 4, 1), (5, 3): (3, 3, 0, 1), (7, 1): (4, 1), (7, 2): (3, 6, 1),
    (11, 1): (9, 1), (11, 2): (2, 7, 1), (13, 1): (11, 1), (13, 2): (2, 12, 1),
}


def _prime_power(q: int) -> Tuple[int, int]:
    for p in range(2, q + 1):
        if q % p == 0:
            k, x = 0, q
            while x % p == 0:
                x //= p
                k += 1
            if x != 1:
                break
            return p, k
    raise ValueError(f"{q} is not a prime power")


class Field:
    """GF(p^k), elements labelled 0..q-1 by the base-p digits of their polynomial (constant term
    lowest), arithmetic by table lookup (vectorised over numpy index arrays)."""

    def __init__(self, order: int):
        p, k = _prime_power(int(order))
        self.order, self.characteristic, self.degree = int(order), p, k
        q = self.order
        digits = np.empty((q, k), dtype=np.int64)
        x = np.arange(q)
        for d in range(k):
            digits[:, d] = x % p
            x = x // p
        weights = p ** np.arange(k)
        self.add = (((digits[:, None, :] + digits[None, :, :]) % p) @ weights).astype(np.int64)
        self.neg = (((-digits) % p) @ weights).astype(np.int64)
        if k == 1:
            self.mul = (np.arange(q)[:, None] * np.arange(q)[None, :]) % p
            self.modulus = (0, 1)
        else:
            self.modulus = _CONWAY.get((p, k)) or self._find_modulus(p, k)
            self.mul = self._mul_table(digits, weights, p, k, self.modulus)
        self.inv = np.zeros(q, dtype=np.int64)            # inv[0] stays 0 (never used)
        rows, cols = np.nonzero(self.mul == 1)
        self.inv[rows] = cols
        if not (np.sort(rows) == np.arange(1, q)).all():
            raise ValueError("modulus is not irreducible")
        self.primitive_element = self._primitive()

    @staticmethod
    def _mul_table(digits, weights, p, k, modulus):
        q = digits.shape[0]
        prod = np.zeros((q, q, 2 * k - 1), dtype=np.int64)
        for i in range(k):
            for j in range(k):
                prod[:, :, i + j] += digits[:, None, i] * digits[None, :, j]
        prod %= p
        for d in range(2 * k - 2, k - 1, -1):             # x^d = -(modulus without its leading term) x^(d-k)
            c = prod[:, :, d].copy()
            for t in range(k):
                prod[:, :, d - k + t] = (prod[:, :, d - k + t] - c * modulus[t]) % p
            prod[:, :, d] = 0
        return (prod[:, :, :k] @ weights).astype(np.int64)

    @classmethod
    def _find_modulus(cls, p, k):
        q = p ** k
        digits = np.empty((q, k), dtype=np.int64)
        x = np.arange(q)
        for d in range(k):
            digits[:, d] = x % p
            x = x // p
        weights = p ** np.arange(k)
        for low in itertools.product(range(p), repeat=k):
            modulus = tuple(low) + (1,)
            mul = cls._mul_table(digits, weights, p, k, modulus)
            if ((mul[1:, 1:] == 1).sum(axis=1) == 1).all() and (mul[1:, 1:] != 0).all():
                # irreducible; prefer one with x primitive
                e, seen = p, 1
                while e != 1:
                    e = int(mul[e, p])
                    seen += 1
                if seen == q - 1:
                    return modulus
        raise ValueError("no primitive polynomial found")

    def _primitive(self) -> int:
        q = self.order
        for g in range(1, q):
            e, n = g, 1
            while e != 1:
                e = int(self.mul[e, g])
                n += 1
            if n == q - 1:
                return g
        return 1                                          # GF(2)

    @property
    def elements(self) -> range:
        return range(self.order)

    def sub(self, a, b):
        return self.add[a, self.neg[b]]


# ---------------------------------------------------------------------------------------------
# finite groups as Cayley tables
# ---------------------------------------------------------------------------------------------

class FiniteGroup:
    """A finite group by its Cayley table ``table[a, b] = index of a*b`` (index 0 = identity) plus a
    label per element (what the reference's ``Group`` objects carry: the canonical 2x2 matrix of a
    PGL(2,q) element, the vector of a Z_q^m element)."""

    def __init__(self, table: np.ndarray, labels: np.ndarray):
        self.table = np.asarray(table, dtype=np.int64)
        self.labels = np.asarray(labels)
        n = self.table.shape[0]
        if not (self.table[0] == np.arange(n)).all() or not (self.table[:, 0] == np.arange(n)).all():
            raise ValueError("element 0 must be the identity")
        self.inverse = np.empty(n, dtype=np.int64)
        rows, cols = np.nonzero(self.table == 0)
        self.inverse[rows] = cols
        self._index = {tuple(np.atleast_1d(l).ravel().tolist()): i for i, l in enumerate(self.labels)}

    @property
    def order(self) -> int:
        return self.table.shape[0]

    def index(self, label) -> int:
        return self._index[tuple(np.asarray(label).ravel().tolist())]

    def generated_by(self, generators: Sequence[int]) -> np.ndarray:
        """Indices reachable from the identity by right multiplication with the generators
        (``_dfs_generators``, ``lifted_product_code.py:201-219``)."""
        seen = np.zeros(self.order, dtype=bool)
        seen[0] = True
        frontier = np.array([0])
        gens = np.asarray(list(generators), dtype=np.int64)
        while frontier.size:
            nxt = np.unique(self.table[frontier[:, None], gens[None, :]])
            nxt = nxt[~seen[nxt]]
            seen[nxt] = True
            frontier = nxt
        return np.flatnonzero(seen)

    def subgroup(self, members: np.ndarray) -> Tuple["FiniteGroup", np.ndarray]:
        """The subgroup on ``members`` (closed, containing index 0) and the old -> new index map."""
        members = np.asarray(members)
        remap = -np.ones(self.order, dtype=np.int64)
        remap[members] = np.arange(members.size)
        table = remap[self.table[np.ix_(members, members)]]
        if (table < 0).any():
            raise ValueError("not closed under multiplication")
        return FiniteGroup(table, self.labels[members]), remap


def _pgl2_canonical(field: Field, m: np.ndarray) -> np.ndarray:
    """Scale (.., 4) matrices [a b; c d] so that a = 1, or b = 1 when a = 0
    (``PGL2.canonicalize``, ``lifted_product_code.py:87-91``)."""
    lead = np.where(m[..., 0] != 0, m[..., 0], m[..., 1])
    s = field.inv[lead]
    return field.mul[m, s[..., None]]


def _mat2_mul(field: Field, x: np.ndarray, y: np.ndarray) -> np.ndarray:
    a, b, c, d = (x[..., i] for i in range(4))
    e, f, g, h = (y[..., i] for i in range(4))
    mul, add = field.mul, field.add
    return np.stack([add[mul[a, e], mul[b, g]], add[mul[a, f], mul[b, h]],
                     add[mul[c, e], mul[d, g]], add[mul[c, f], mul[d, h]]], axis=-1)


def _mat2_det(field: Field, x: np.ndarray) -> np.ndarray:
    return field.sub(field.mul[x[..., 0], x[..., 3]], field.mul[x[..., 1], x[..., 2]])


def pgl2_group(q: int, field: Optional[Field] = None) -> FiniteGroup:
    """PGL(2,q): every invertible 2x2 matrix over GF(q) up to scaling, as canonical representatives
    ``(1, b, c, d)`` / ``(0, 1, c, d)``; order (q-1) q (q+1).  Labels are the flattened matrices."""
    field = field or Field(q)
    q = field.order
    r = np.arange(q)
    first = np.stack(np.meshgrid([1], r, r, r, indexing="ij"), axis=-1).reshape(-1, 4)
    second = np.stack(np.meshgrid([0], [1], r, r, indexing="ij"), axis=-1).reshape(-1, 4)
    mats = np.concatenate([first, second])
    mats = mats[_mat2_det(field, mats) != 0]
    ident = np.flatnonzero((mats == np.array([1, 0, 0, 1])).all(axis=1))[0]
    mats[[0, ident]] = mats[[ident, 0]]
    code = lambda m: ((m[..., 0] * q + m[..., 1]) * q + m[..., 2]) * q + m[..., 3]
    lookup = -np.ones(q ** 4, dtype=np.int64)
    lookup[code(mats)] = np.arange(mats.shape[0])
    n = mats.shape[0]
    table = np.empty((n, n), dtype=np.int64)
    for i in range(n):
        table[i] = lookup[code(_pgl2_canonical(field, _mat2_mul(field, mats[i][None, :], mats)))]
    g = FiniteGroup(table, mats)
    g.field = field
    return g


def psl2_group(q: int) -> FiniteGroup:
    """PSL(2,q) inside PGL(2,q): the classes with a representative of determinant 1, i.e. whose
    determinant is a square (``get_psl2``, ``lifted_product_code.py:192-199``)."""
    pgl = pgl2_group(q)
    f = pgl.field
    squares = np.zeros(f.order, dtype=bool)
    squares[f.mul[np.arange(f.order), np.arange(f.order)]] = True
    members = np.flatnonzero(squares[_mat2_det(f, pgl.labels)])
    sub, _ = pgl.subgroup(members)
    sub.field = f
    return sub


def get_psl2(q: int) -> frozenset:
    """Canonical matrices (as 4-tuples ``(a, b, c, d)`` of field labels) of all elements of PSL(2,q)."""
    return frozenset(tuple(int(x) for x in m) for m in psl2_group(q).labels)


def zqm_group(q: int, m: int) -> FiniteGroup:
    """The abelian group Z_q^m (``Zqm``, ``lifted_product_code.py:104-137``); labels are the vectors."""
    vecs = np.array(list(itertools.product(range(q), repeat=m)), dtype=np.int64).reshape(-1, m)
    weights = q ** np.arange(m - 1, -1, -1)
    s = (vecs[:, None, :] + vecs[None, :, :]) % q
    return FiniteGroup(s @ weights, vecs)


# ---------------------------------------------------------------------------------------------
# generators and local codes
# ---------------------------------------------------------------------------------------------

def random_check_matrix(r: int, n: int, seed=None, full_rank: Optional[bool] = None) -> np.ndarray:
    """Random r x n matrix over GF(2) (``random_code.py:4-23``; same numpy draws for a seed)."""
    rng = np.random.default_rng(seed)
    for _ in range(10000):
        h = rng.integers(low=0, high=2, size=(r, n)).astype(np.uint8)
        if not full_rank or gf2.rank(h) == min(h.shape):
            return h
    raise RuntimeError("Failed to construct random matrix: Number of retries exceeded")


def random_abelian_generators(q: int, m: int, k: int, symmetric: Optional[bool] = None, seed=None) -> np.ndarray:
    """k random elements of Z_q^m as the rows of a ``(k, m)`` array (not necessarily generating or
    independent); ``symmetric``: k/2 draws, each followed by its inverse
    (``lifted_product_code.py:139-157``; same numpy draws for a seed)."""
    rng = np.random.default_rng(seed)
    symmetrize = bool(symmetric) and q != 2
    if symmetrize:
        if k % 2 != 0:
            raise ValueError("Number of generators must be even when the set is symmetrized and q /= 2")
        k //= 2
    gens = rng.integers(low=0, high=q, size=(k, m)).astype(np.int64)
    if symmetrize:
        gens = np.stack([gens, (q - gens) % q], axis=1).reshape(2 * k, m)
    return gens


def morgenstern_matrices(l: int, i: int) -> Tuple[Field, np.ndarray]:
    """The q + 1 Morgenstern matrices ``[[1, g + d i], [x (g + d + d i), 1]]`` of PGL(2, q^i),
    q = 2^l, over the (g, d) in F_q^2 with g^2 + g d + d^2 eps = 1 (``lifted_product_code.py:159-185``,
    after Dinur et al. arXiv:2111.04808): returns the field GF(q^i) and the ``(q+1, 4)`` matrices."""
    if l < 1:
        raise ValueError("l must be positive")
    if i % 2 != 0:
        raise ValueError("The Morgenstern construction works only for PGL(2, q^i) with even i, "
                         "because we need a quaternion algebra")
    q = 2 ** l
    big = Field(q ** i)
    mul, add = big.mul, big.add
    # F_q inside GF(q^i): the roots of x^q = x.  The reference takes "labels < q" for F_q, which is
    # the subfield only for q = 2 (for larger q a polynomial of degree < l in the generator of
    # GF(q^i) is in general not in F_q); here the true subfield is used.
    def power(x, e):
        r = np.ones_like(x)
        for _ in range(e):
            r = mul[r, x]
        return r
    allx = np.arange(big.order)
    in_fq = power(allx, q) == allx
    fq = allx[in_fq]
    # an element outside F_q with i^2 + i in F_q
    cand = [int(x) for x in allx if not in_fq[x] and in_fq[add[mul[x, x], x]]]
    if not cand:
        raise ValueError("no element i with i^2 + i in F_q")
    i_el = cand[0]
    eps = int(add[mul[i_el, i_el], i_el])
    pairs = [(int(g), int(d)) for g in fq for d in fq
             if add[add[mul[g, g], mul[g, d]], mul[mul[d, d], eps]] == 1]
    if len(pairs) != q + 1:
        raise AssertionError(f"expected {q + 1} solutions, found {len(pairs)}")
    x = big.primitive_element
    mats = np.array([[1, add[g, mul[d, i_el]], mul[x, add[add[g, d], mul[d, i_el]]], 1] for g, d in pairs], dtype=np.int64)
    return big, mats


def morgenstern_generators(l: int, i: int, use_B_generators: Optional[bool] = None,
                           symmetric: Optional[bool] = None) -> Tuple[FiniteGroup, List[int]]:
    """PGL(2, q^i) and the indices of the Morgenstern generators in it; with ``use_B_generators``
    the products ``a b`` of distinct generators (ordered pairs when ``symmetric``, the default, as
    ``(a b)^-1 = b a``) (``lifted_product_code.py:159-190``)."""
    symmetric = True if symmetric is None else symmetric
    field, mats = morgenstern_matrices(l, i)
    group = pgl2_group(field.order, field)
    gens = [group.index(_pgl2_canonical(field, m)) for m in mats]
    if use_B_generators:
        gens = [int(group.table[a, b]) for ia, a in enumerate(gens) for ib, b in enumerate(gens)
                if ia != ib and (ia < ib or symmetric)]
    return group, gens


# ---------------------------------------------------------------------------------------------
# the lifted product over a Cayley graph
# ---------------------------------------------------------------------------------------------

def lifted_product_code(group: FiniteGroup, generators: Sequence[int], h1, h2, check_complex: Optional[bool] = None,
                        compute_logicals: Optional[bool] = None, double_cover: Optional[bool] = None) -> QuantumCode:
    """E x V -> E x E + V x V -> V x E over the Cayley graph of ``generators`` in ``group`` with local
    codes ``h1`` (left factor, group acting from the left) and ``h2`` (right factor, from the right);
    ``double_cover`` (default) takes the base graph D_w, else the bouquet B_w whose local codes have
    length 2w: out-edges first, then in-edges (``lifted_product_code.py:249-404``)."""
    warnings.warn("Lifted Product codes is experimental!")
    double_cover = True if double_cover is None else double_cover
    h1 = np.asarray(h1) % 2
    h2 = np.asarray(h2) % 2
    if h1.shape[1] != h2.shape[1]:
        raise ValueError("Local code block lengths must match. (For now)")
    gens = [int(g) for g in generators]
    w = len(gens)
    if h1.shape[1] != (w if double_cover else 2 * w):
        raise ValueError("Local code block length does not match base graph degree")
    # base graph: edge t carries generator t; (tail, head, local column at the tail, at the head)
    if double_cover:
        verts = 2
        edges = [(0, 1, t, t) for t in range(w)]
    else:
        verts = 1
        edges = [(0, 0, t, w + t) for t in range(w)]

    def boundary(h, head_elem):
        rows = [[[] for _ in range(w)] for _ in range(verts * h.shape[0])]
        for t, (tail, head, out_col, in_col) in enumerate(edges):
            for r in range(h.shape[0]):
                if h[r, in_col]:
                    rows[head * h.shape[0] + r][t].append(head_elem[t])
                if h[r, out_col]:
                    rows[tail * h.shape[0] + r][t].append(0)
        return rows

    a = boundary(h1, gens)                                               # (V r1) x E, left action
    bt = boundary(h2, [int(group.inverse[g]) for g in gens])             # (V r2) x E ...
    b = [[bt[i][j] for i in range(len(bt))] for j in range(w)]           # ... transposed: E x (V r2), right action
    code = lifted_product(group.table, a, b, compute_logicals=bool(compute_logicals))
    if check_complex:
        hx, hz = code.checks.x, code.checks.z
        if ((hz @ hx.T).data % 2).any():
            raise AssertionError("boundary maps do not compose to zero")
        if compute_logicals:
            lx, lz = code.logicals.x.astype(np.int64), code.logicals.z.astype(np.int64)
            assert not ((hz @ lx.T) % 2).any() and not ((hx @ lz.T) % 2).any()
            assert lx.shape[0] + gf2.rank(hx.toarray()) + gf2.rank(hz.toarray()) == code.num_qubits
    return code


def _wrapper(group: FiniteGroup, gens, r, compute_logicals, seed, check_complex, r2=None, double_cover=None):
    """``_lifted_product_code_wrapper`` (``lifted_product_code.py:406-422``): the group generated by
    the generators, local codes drawn with seeds ``seed + 1`` and ``seed + 2``."""
    if r <= 0:
        raise ValueError("r must be positive")
    r2 = r if r2 is None else r2
    compute_logicals = True if compute_logicals is None else compute_logicals
    w = len(gens)
    members = group.generated_by(gens)
    if members.size != group.order:
        group, remap = group.subgroup(members)
        gens = [int(remap[g]) for g in gens]
    width = w if double_cover else 2 * w
    h1 = random_check_matrix(r, width, seed=seed + 1 if seed is not None else None)
    h2 = random_check_matrix(r2, width, seed=seed + 2 if seed is not None else None)
    return lifted_product_code(group, gens, h1, h2, check_complex=check_complex, compute_logicals=compute_logicals,
                               double_cover=double_cover)


def lifted_product_code_cyclic(q, m, w, r, compute_logicals=None, r2=None, seed=None, check_complex=None,
                               double_cover=None) -> QuantumCode:
    """Lifted product with w random generators of Z_q^m (``lifted_product_code.py:424-439``;
    ``double_cover`` defaults to False here, as there)."""
    if q <= 0 or m <= 0 or w <= 0:
        raise ValueError("q, m and w must be positive")
    double_cover = False if double_cover is None else double_cover
    group = zqm_group(q, m)
    gens = [group.index(v) for v in random_abelian_generators(q, m, w, seed=seed)]
    return _wrapper(group, gens, r, compute_logicals, seed, check_complex, r2=r2, double_cover=double_cover)


def lifted_product_code_pgl2(l, i, r, compute_logicals=None, seed=None, check_complex=None, r2=None,
                             double_cover=None) -> QuantumCode:
    """Lifted product over PGL(2, (2^l)^i) with the Morgenstern generators
    (``lifted_product_code.py:441-454``).  ``double_cover=None`` means True here; the reference sizes
    the local codes by the truthiness of the argument and then defaults the base graph to D_w, so it
    raises on ``None`` — its tests always pass the flag."""
    group, gens = morgenstern_generators(l, i)
    double_cover = True if double_cover is None else double_cover
    return _wrapper(group, gens, r, compute_logicals, seed, check_complex, r2=r2, double_cover=double_cover)


def psl2_tanner_lp_4320(compute_logicals: bool = True) -> QuantumCode:
    """[[4320, 28]]: a decodable instance of BASELINE.json configs[3] — the lifted product over the
    Cayley graph of PGL(2,4) = PSL(2,4) ~ A5 (order 60) on the six Morgenstern "B" generators
    ``{a b : a != b}`` (double cover D_6), both local codes the [6,3,3] shortened Hamming code.
    n = (6^2 + 4 * 3^2) * 60; X and Z checks of weight 5 / 7, qubits in 2..4 checks of either type.
    (``lifted_product_code_pgl2(1, 2, 5)`` of the reference's tests draws 5 x 3 local codes at random:
    all-zero checks and untouched qubits — it pins the construction, not a code one would decode.)"""
    local = np.array([[1, 1, 0, 1, 0, 0], [0, 1, 1, 0, 1, 0], [1, 0, 1, 0, 0, 1]], dtype=np.uint8)
    group, gens = morgenstern_generators(1, 2, use_B_generators=True, symmetric=True)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        return lifted_product_code(group, gens, local, local, compute_logicals=compute_logicals, double_cover=True)

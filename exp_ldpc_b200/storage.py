"""Storage-simulation circuits and their noise models (set-up time; host only).

Host mirror of what the reference hands to Stim: ``build_storage_simulation``
(``python/qldpc/storage_sim.py:110-199``), the syndrome-extraction round it repeats
(``build_perfect_circuit``, ``:38-75``), and the text-rewriting noise models ``trivial_noise`` /
``depolarizing_noise`` / ``circuit_noise`` (``python/qldpc/noise_model.py:11-13,117-151``).  Same
function names, arguments, qubit numbering (data ``0..n-1``, X ancillas, then Z ancillas), record
layout and instruction formats, so the emitted text feeds ``exp_ldpc_b200.circuits`` exactly like
the reference-generated fixtures do.

One deliberate difference: the CX / CZ layers come from this module's own bipartite edge colouring
(:func:`schedule_checks`), not from a transcription of the reference's Kempe-chain routine, whose
output order depends on Python ``set`` iteration.  Both give a proper colouring with max-degree many
layers (what ``tests/test_edge_coloring.py`` of the reference checks); which gate sits in which layer
may differ, the measurement record layout and — for noise that only acts at measurement steps —
the sampled distribution do not.
"""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Callable, Dict, Iterable, List, Optional, Sequence

import numpy as np

from .codes import QuantumCode, canonical_csr

__all__ = ["CircuitTargets", "NoiseRewriter", "StorageSim", "schedule_checks", "edge_color_bipartite", "build_perfect_circuit",
           "build_storage_simulation", "trivial_noise", "depolarizing_noise", "circuit_noise", "circuit_ticks"]


@dataclass(frozen=True)
class CircuitTargets:            # qecc_util.py:120-131
    data: List[int]
    x_checks: List[int]
    z_checks: List[int]

    @property
    def ancillas(self) -> List[int]:
        return list(self.x_checks) + list(self.z_checks)


@dataclass(frozen=True)
class StorageSim:                # qecc_util.py:151-155
    circuit: List[str]
    measurement_view: Callable
    data_view: Callable


@dataclass(frozen=True)
class NoiseRewriter:
    """``rewrite(targets, circuit) -> lines`` like the reference's ``NoiseRewriter``; ``kind``/``p``/``pm``
    let the rounds = 0 fast path use the closed-form flip probability instead of a circuit."""
    rewrite: Callable
    kind: str = "custom"
    p: float = 0.0
    pm: float = 0.0

    def code_capacity_flip_probability(self) -> float:
        if self.kind == "trivial":
            return 0.0
        if self.kind != "depolarizing":
            raise NotImplementedError("closed-form code-capacity rate exists for depolarizing_noise only")
        a = 2.0 * self.p / 3.0                      # X or Y from DEPOLARIZE1(p), then an independent readout flip
        return a + self.pm - 2.0 * a * self.pm


# ---------------------------------------------------------------------------------------------
# scheduling: proper edge colouring of the Tanner graph with max-degree colours
# ---------------------------------------------------------------------------------------------
def schedule_checks(checks) -> List[Dict[int, int]]:
    """``layers[c][check] = data qubit``: every check touches each of its qubits in exactly one layer and
    no qubit or check appears twice in a layer; ``len(layers)`` = maximum degree (Koenig).
    Alternating-path colouring over the CSR edge list, deterministic."""
    h = canonical_csr(checks)
    m, n = h.shape
    deg = 0
    if h.nnz:
        deg = int(max(np.diff(h.indptr).max(), np.bincount(h.indices, minlength=n).max()))
    at_check = -np.ones((m, deg), dtype=np.int64)    # at_check[c, colour] = data qubit
    at_data = -np.ones((n, deg), dtype=np.int64)     # at_data[q, colour] = check
    for c in range(m):
        for q in h.indices[h.indptr[c]:h.indptr[c + 1]]:
            a = int(np.flatnonzero(at_check[c] < 0)[0])
            b = int(np.flatnonzero(at_data[q] < 0)[0])
            if a != b and at_data[q, a] >= 0:
                # free colour a at q: flip a <-> b along the alternating path that leaves q by its a-edge
                path = []
                on_data, v, col = True, int(q), a
                while True:
                    w = int(at_data[v, col]) if on_data else int(at_check[v, col])
                    if w < 0:
                        break
                    path.append((w, v, col) if on_data else (v, w, col))       # (check, data, colour)
                    v, on_data, col = w, not on_data, (b if col == a else a)
                for cc, qq, col in path:
                    at_check[cc, col] = -1
                    at_data[qq, col] = -1
                for cc, qq, col in path:
                    new = b if col == a else a
                    at_check[cc, new] = qq
                    at_data[qq, new] = cc
            at_check[c, a] = q
            at_data[q, a] = c
    return [{int(c): int(at_check[c, col]) for c in range(m) if at_check[c, col] >= 0} for col in range(deg)]


def edge_color_bipartite(biadjacency) -> List[set]:
    """Proper edge colouring of a bipartite (multi)graph with maximum-degree many colours (Koenig): the counterpart
    of the reference's ``edge_color_bipartite`` (``edge_coloring.py:17-88``) on a biadjacency matrix instead of a
    networkx graph.  ``biadjacency[u, v]`` = number of parallel edges between left vertex u and right vertex v;
    returns one set per colour of edges ``(u, v, key)``, ``key`` numbering the parallel edges of a pair.
    Same alternating-path argument as :func:`schedule_checks`, kept separate because that one fixes the gate order
    of the storage circuits."""
    a = np.asarray(biadjacency.toarray() if hasattr(biadjacency, "toarray") else biadjacency).astype(np.int64)
    if a.ndim != 2 or (a < 0).any():
        raise RuntimeError("biadjacency must be a matrix of non-negative edge multiplicities")
    nl, nr = a.shape
    deg = int(max(a.sum(axis=1).max(initial=0), a.sum(axis=0).max(initial=0)))
    edges = [(u, v, k) for u in range(nl) for v in np.flatnonzero(a[u]) for k in range(a[u, v])]
    colour = -np.ones(len(edges), dtype=np.int64)
    at_left = -np.ones((nl, deg), dtype=np.int64)       # at_left[u, colour] = edge id
    at_right = -np.ones((nr, deg), dtype=np.int64)
    for e, (u, v, _) in enumerate(edges):
        ca = int(np.flatnonzero(at_left[u] < 0)[0])     # free at u
        cb = int(np.flatnonzero(at_right[v] < 0)[0])    # free at v
        if ca != cb and at_right[v, ca] >= 0:
            # walk the ca / cb alternating path that leaves v by its ca-edge (it cannot reach u) and swap the colours
            path, on_right, x, col = [], True, int(v), ca
            while True:
                f = int(at_right[x, col]) if on_right else int(at_left[x, col])
                if f < 0:
                    break
                path.append(f)
                x = edges[f][0] if on_right else edges[f][1]
                on_right, col = not on_right, (cb if col == ca else ca)
            for f in path:
                at_left[edges[f][0], colour[f]] = -1
                at_right[edges[f][1], colour[f]] = -1
            for f in path:
                colour[f] = cb if colour[f] == ca else ca
                at_left[edges[f][0], colour[f]] = f
                at_right[edges[f][1], colour[f]] = f
        colour[e] = ca
        at_left[u, ca] = e
        at_right[v, ca] = e
    return [{edges[e] for e in np.flatnonzero(colour == c)} for c in range(deg)]


def build_perfect_circuit(code: QuantumCode):
    """One noiseless syndrome-extraction round: X checks (RX, CX layers, MRX), then Z checks (RX, CZ
    layers, MRX); no trailing TICK (``storage_sim.py:38-75``)."""
    n = code.num_qubits
    mx, mz = code.checks.x.shape[0], code.checks.z.shape[0]
    x_anc = list(range(n, n + mx))
    z_anc = list(range(n + mx, n + mx + mz))
    join = lambda v: " ".join(str(i) for i in v)                              # noqa: E731
    lines = [f"RX {join(x_anc)}", "TICK"]
    if mx:
        for layer in schedule_checks(code.checks.x):
            lines += [f"CX {x_anc[c]} {q}" for c, q in layer.items()]
            lines.append("TICK")
        lines.append(f"MRX {join(x_anc)}")
    lines += [f"RX {join(z_anc)}", "TICK"]
    if mz:
        for layer in schedule_checks(code.checks.z):
            lines += [f"CZ {z_anc[c]} {q}" for c, q in layer.items()]
            lines.append("TICK")
        lines.append(f"MRX {join(z_anc)}")
    return CircuitTargets(list(range(n)), x_anc, z_anc), lines


def build_storage_simulation(rounds: int, noise_model: NoiseRewriter, code: QuantumCode, use_x_logicals=None) -> StorageSim:
    """Prepare logical 0 (or + with ``use_x_logicals``), run ``rounds`` rounds of syndrome extraction,
    read the data out transversally (``storage_sim.py:110-199``).  Record layout per shot:
    ``[round 0: X(mx) | Z(mz)] ... [round T-1] [data(n)]``."""
    use_x = bool(use_x_logicals)
    basis = "X" if use_x else "Z"
    targets, round_lines = build_perfect_circuit(code)
    n, mx, mz = len(targets.data), len(targets.x_checks), len(targets.z_checks)
    per_round = mx + mz
    data_str = " ".join(str(i) for i in targets.data)
    out: List[str] = [f"R{basis} {data_str}", "TICK"]
    if rounds > 0:
        out += round_lines
        first = range(0, mx) if use_x else range(mx, per_round)               # only one type is deterministic at first
        out += [f"DETECTOR(0, {i}) rec[{i - per_round}]" for i in first]
        if rounds > 1:
            out += ["TICK", f"REPEAT {rounds - 1} {{"]
            out += round_lines
            out.append("SHIFT_COORDS(1, 0)")
            out += [f"DETECTOR(0, {i}) rec[{i - per_round}] rec[{i - 2 * per_round}]" for i in range(per_round)]
            out += ["TICK", "}"]
    out.append(f"M{basis} {data_str}")
    recs = lambda support: " ".join(f"rec[{int(v) - n}]" for v in support)     # noqa: E731
    checks = canonical_csr(code.checks.x if use_x else code.checks.z)
    logicals = np.asarray(code.logicals.x if use_x else code.logicals.z)
    prev_shift = 0 if use_x else mx
    for i in range(checks.shape[0]):
        prev = f"rec[{i - n - per_round + prev_shift}] " if rounds > 0 else ""
        out.append(f"DETECTOR(1, {i}) " + prev + recs(checks.indices[checks.indptr[i]:checks.indptr[i + 1]]))
    for i in range(logicals.shape[0]):
        out.append(f"OBSERVABLE_INCLUDE({i}) " + recs(np.flatnonzero(logicals[i])))
    out = list(noise_model.rewrite(targets, out))

    def measurement_view(round_index, get_x_checks, measurement_vector, *_):
        off = per_round * round_index + (0 if get_x_checks else mx)
        return measurement_vector[off:off + (mx if get_x_checks else mz)]

    def data_view(measurement_vector, *_):
        return measurement_vector[per_round * rounds:per_round * rounds + n]

    return StorageSim(out, measurement_view, data_view)


# ---------------------------------------------------------------------------------------------
# noise models: text rewriting (noise_model.py)
# ---------------------------------------------------------------------------------------------
_MEASURE = re.compile(r"^(?:\s*)(M|MZ|MX|MY|MPP|MR|MRZ|MRX|MRY)((?:\s*\d+\s*)+)$")
_TWO_QUBIT = {"CNOT", "CX", "CY", "CZ", "ISWAP", "ISWAP_DAG", "SQRT_XX", "SQRT_XX_DAG", "SQRT_YY", "SQRT_YY_DAG", "SQRT_ZZ",
              "SQRT_ZZ_DAG", "SWAP", "XCX", "XCY", "XCZ", "YCX", "YCY", "YCZ", "ZCX", "ZCY", "ZCZ"}


def _tokens(line: str) -> List[str]:
    return [t.upper() for t in line.split("#")[0].split()]


def circuit_ticks(circuit: Iterable[str]) -> List[List[str]]:
    """Split into time steps; a ``TICK`` line opens the next step (``noise_model.py:30-67``)."""
    steps: List[List[str]] = [[]]
    for line in circuit:
        tk = _tokens(line)
        if tk and tk[0] == "TICK":
            steps.append([])
        steps[-1].append(line)
    return steps


def _noisy_measurement(line: str, pm) -> str:
    m = _MEASURE.search(line)
    return line if m is None else f"{m.group(1)}({pm}){m.group(2)}"


def trivial_noise() -> NoiseRewriter:
    return NoiseRewriter(lambda targets, circuit: list(circuit), "trivial")


def depolarizing_noise(p: float, pm: float) -> NoiseRewriter:
    """``DEPOLARIZE1(p)`` on the data qubits at the start of every time step that contains a measurement;
    measurements flip with probability ``pm`` (``noise_model.py:117-123``)."""
    def rewrite(targets: CircuitTargets, circuit: Sequence[str]) -> List[str]:
        noise = f"DEPOLARIZE1({p}) " + " ".join(str(i) for i in targets.data)
        out: List[str] = []
        for step in circuit_ticks(circuit):
            if not any(_MEASURE.search(ln) for ln in step):
                out += step
                continue
            body = step
            if _tokens(step[0])[:1] == ["TICK"]:
                out.append(step[0])
                body = step[1:]
            out.append(noise)
            out += [_noisy_measurement(ln, pm) for ln in body]
        return out
    return NoiseRewriter(rewrite, "depolarizing", float(p), float(pm))


def circuit_noise(p: float, pm: Optional[float] = None) -> NoiseRewriter:
    """After every time step: ``DEPOLARIZE2(p)`` on the pairs of its two-qubit gates, ``DEPOLARIZE1(p)`` on
    every other qubit; measurements flip with probability ``pm`` (default ``p``)
    (``noise_model.py:125-151``)."""
    pm = p if pm is None else pm

    def rewrite(targets: CircuitTargets, circuit: Sequence[str]) -> List[str]:
        everyone = set(targets.data) | set(targets.ancillas)
        out: List[str] = []
        for step in circuit_ticks(circuit):
            pairs = []
            for ln in step:
                tk = _tokens(ln)
                if len(tk) > 1 and tk[0] in _TWO_QUBIT:
                    q = [int(v) for v in tk[1:]]
                    if len(q) % 2:
                        raise ValueError(f"Found an odd number of targets for a two qubit gate directive: \n f{ln}")
                    pairs += list(zip(q[::2], q[1::2]))
            out += [_noisy_measurement(ln, pm) for ln in step]
            if pairs:
                out.append(f"DEPOLARIZE2({p}) " + " ".join(f"{a} {b}" for a, b in pairs))
            idle = sorted(everyone - {q for pr in pairs for q in pr})
            out.append(f"DEPOLARIZE1({p}) " + " ".join(str(q) for q in idle))
        return out
    return NoiseRewriter(rewrite, "circuit", float(p), float(pm))

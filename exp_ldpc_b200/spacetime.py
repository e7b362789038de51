"""Decoding-problem builders: the matrices the decoder sees and the per-shot pre/post steps.

Host-side mirror of ``python/qldpc/spacetime_code.py`` (names and semantics kept):

* :class:`SpacetimeCode` — ``H_st = [ blockdiag(H x (T+1)) | M ]`` (``spacetime_code.py:46-75``);
* :class:`SpacetimeCodeSingleShot` — ``[H | I_m]`` (``spacetime_code.py:10-37``);
* :func:`spacetime_syndrome` — history + transversal readout -> differenced syndrome
  (``spacetime_code.py:98-119``);
* ``final_correction`` — XOR-fold of the ``T+1`` data blocks (``spacetime_code.py:81-84``);
* :class:`DetectorSpacetimeCode` — fault check matrix, fault map and priors of a detector error
  model given as Stim ``.dem`` text (``spacetime_code.py:122-183``).

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

import re
from typing import Callable

import numpy as np
from scipy import sparse

from .codes import canonical_csr

__all__ = ["SpacetimeCode", "SpacetimeCodeSingleShot", "DetectorSpacetimeCode", "spacetime_syndrome", "MeasurementLayout",
           "flatten_detector_model", "detector_model_text"]


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


# ---------------------------------------------------------------------------------------------
# detector error models in Stim's text format
# ---------------------------------------------------------------------------------------------

def flatten_detector_model(text) -> list:
    """``[(p, detectors, observables), ...]`` plus declared ids of a detector error model in Stim's
    ``.dem`` text: ``error(p) D.. L.. [^ ..]``, ``detector[(coords)] D..``, ``logical_observable L..``,
    ``shift_detectors[(coords)] k`` and ``repeat k { .. }`` blocks — what ``DetectorErrorModel.flattened()``
    hands the reference (``spacetime_code.py:131``): relative detector ids made absolute, loops unrolled.
    Returns ``(errors, max_detector_id, max_observable_id)`` (ids -1 when none is named)."""
    lines = text.splitlines() if isinstance(text, str) else [str(ln) for ln in text]
    toks = []
    for ln in lines:
        ln = ln.split("#", 1)[0].strip()
        if ln:
            toks.append(ln)
    errors, state = [], {"offset": 0, "max_d": -1, "max_l": -1}

    def block_end(start):
        depth = 0
        for k in range(start, len(toks)):
            depth += toks[k].endswith("{")
            if toks[k] == "}":
                depth -= 1
                if depth == 0:
                    return k
        raise ValueError("unterminated repeat block in detector error model")

    def run(lo, hi):
        k = lo
        while k < hi:
            ln = toks[k]
            mt = re.match(r"^(\w+)\s*(?:\(([^)]*)\))?\s*(.*)$", ln)
            if mt is None:
                raise ValueError(f"cannot parse detector error model line {ln!r}")
            name, arg, rest = mt.group(1).lower(), mt.group(2) or "", mt.group(3)
            if name == "repeat":
                if not ln.endswith("{"):
                    raise ValueError(f"malformed repeat: {ln!r}")
                end = block_end(k)
                for _ in range(int(rest.rstrip("{").strip())):
                    run(k + 1, end)
                k = end + 1
                continue
            targets = [t for t in rest.split() if t != "^"]
            if name == "shift_detectors":
                state["offset"] += int(targets[0]) if targets else 0
            elif name in ("error", "detector", "logical_observable"):
                dets = [int(t[1:]) + state["offset"] for t in targets if t[0] in "dD"]
                obs = [int(t[1:]) for t in targets if t[0] in "lL"]
                if len(dets) + len(obs) != len(targets):
                    raise ValueError(f"unknown target in {ln!r}")
                state["max_d"] = max([state["max_d"]] + dets)
                state["max_l"] = max([state["max_l"]] + obs)
                if name == "error":
                    args = [a for a in arg.split(",") if a.strip()]
                    if len(args) != 1:
                        raise ValueError(f"error instruction needs exactly one probability: {ln!r}")
                    errors.append((float(args[0]), dets, obs))
            else:
                raise ValueError(f"unsupported detector error model instruction {name!r}")
            k += 1

    run(0, len(toks))
    return errors, state["max_d"], state["max_l"]


def detector_model_text(fault_check_matrix, fault_map, fault_priors) -> str:
    """The three fields as Stim ``.dem`` text (one ``error`` line per column; the detector and observable
    counts are declared so that trailing ids no fault touches survive the round trip)."""
    fc, fm = sparse.csc_matrix(fault_check_matrix), sparse.csc_matrix(fault_map)
    out = []
    for j, p in enumerate(np.asarray(fault_priors, dtype=np.float64)):
        t = [f"D{int(d)}" for d in np.sort(fc.indices[fc.indptr[j]:fc.indptr[j + 1]])]
        t += [f"L{int(o)}" for o in np.sort(fm.indices[fm.indptr[j]:fm.indptr[j + 1]])]
        out.append(f"error({float(p)!r}) " + " ".join(t))
    if fc.shape[0]:
        out.append(f"detector D{fc.shape[0] - 1}")
    if fm.shape[0]:
        out.append(f"logical_observable L{fm.shape[0] - 1}")
    return "\n".join(out) + "\n"


class DetectorSpacetimeCode:
    """``fault_check_matrix`` (detectors x faults), ``fault_map`` (observables x faults), both CSR uint32,
    and ``fault_priors`` of a detector error model (``spacetime_code.py:122-183``).  The model comes as
    Stim ``.dem`` text, as any object with a ``flattened()`` whose ``str`` is that text (a
    ``stim.DetectorErrorModel`` where stim is installed), or as :class:`circuits.DetectorErrorModel`.

    Row index = detector / observable id, which is what the construction means and what ``bpd_detector``
    (``_experiment.py:128-151``) needs; the reference's ``lil_to_coo`` writes the position of the id in
    the error's target list instead (``:168-171``), i.e. every error touches rows 0..len-1 — not
    reproduced.  A detector named twice by one error (across ``^`` components) cancels."""

    def __init__(self, detector_model):
        if all(hasattr(detector_model, a) for a in ("fault_check_matrix", "fault_map", "fault_priors")):
            self.fault_check_matrix = sparse.csr_matrix(detector_model.fault_check_matrix).astype(np.uint32)
            self.fault_map = sparse.csr_matrix(detector_model.fault_map).astype(np.uint32)
            self.fault_priors = np.array(detector_model.fault_priors, dtype=np.float64)
            return
        if hasattr(detector_model, "flattened"):
            detector_model = str(detector_model.flattened())
        errors, max_d, max_l = flatten_detector_model(detector_model)

        def incidence(which, rows):
            ij = [(i, j) for j, e in enumerate(errors) for i in e[which]]
            a = sparse.coo_matrix((np.ones(len(ij), dtype=np.int64), ([i for i, _ in ij], [j for _, j in ij])),
                                  shape=(rows, len(errors))).tocsr()
            a.data %= 2
            a.eliminate_zeros()
            a.sort_indices()
            return a.astype(np.uint32)
        self.fault_check_matrix = incidence(1, max_d + 1)
        self.fault_map = incidence(2, max_l + 1)
        self.fault_priors = np.array([e[0] for e in errors], dtype=np.float64)

    @property
    def num_detectors(self) -> int:
        return self.fault_check_matrix.shape[0]

    @property
    def num_faults(self) -> int:
        return self.fault_check_matrix.shape[1]

    def text(self) -> str:
        return detector_model_text(self.fault_check_matrix, self.fault_map, self.fault_priors)

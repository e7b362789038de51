"""Circuit-level sampling: compile the Stim-text circuits of the storage simulation and sample
their measurement records with Pauli frames on the GPU.

The reference builds a list of Stim instruction strings (``build_storage_simulation``,
``python/qldpc/storage_sim.py:110-199``; noise rewriting ``python/qldpc/noise_model.py:117-193``)
and hands it to ``stim.Circuit(...).compile_sampler().sample(samples)``
(``python/qldpc/misc/_experiment.py:171-197``).  :func:`compile_circuit` parses exactly that text
(the instruction subset those builders emit) into the flat program ``qd_circuit_create`` takes;
:class:`CircuitSampler` runs it.  Records come back relative to the noiseless reference sample:
every deterministic outcome of these circuits (Z checks on |0...0>, the data readout) is 0 there,
which is what the decoders consume; X-check outcomes of the first round are random in a real run
and read 0 + frame flips here (the reference never looks at them, ``_experiment.py:207``).
"""
from __future__ import annotations

import ctypes as C
import re
from dataclasses import dataclass
from typing import Iterable, List, Sequence

import numpy as np

from . import _lib

__all__ = ["CircuitProgram", "compile_circuit", "CircuitSampler", "OPCODES", "DetectorErrorModel", "detector_error_model",
           "depolarize1_independent_probability", "depolarize2_independent_probability"]

OPCODES = {"RESET": 0, "CX": 1, "CZ": 2, "MZ": 3, "MX": 4, "MRZ": 5, "MRX": 6, "DEPOLARIZE1": 7, "DEPOLARIZE2": 8,
           "X_ERROR": 9, "Z_ERROR": 10}
_ALIASES = {"R": "RESET", "RZ": "RESET", "RX": "RESET", "M": "MZ", "MR": "MRZ", "CNOT": "CX", "ZCX": "CX", "ZCZ": "CZ"}
_ANNOTATIONS = {"TICK", "DETECTOR", "OBSERVABLE_INCLUDE", "SHIFT_COORDS", "QUBIT_COORDS"}
_LINE = re.compile(r"^\s*([A-Za-z_][A-Za-z_0-9]*)\s*(?:\(([^)]*)\))?\s*(.*?)\s*$")


@dataclass
class CircuitProgram:
    ops: np.ndarray              # (num_ops, 4) int32: opcode | threshold index << 8, q0, q1, noise site
    thresholds: np.ndarray       # uint32, entry 0 = 0
    num_qubits: int
    num_measurements: int
    num_sites: int
    detectors: list = None       # per DETECTOR annotation: absolute measurement indices it XORs
    observables: list = None     # per logical observable id: absolute measurement indices


def _unroll(lines: Sequence[str]) -> List[str]:
    """Expand ``REPEAT n { ... }`` blocks (nesting allowed)."""
    out: List[str] = []
    stack: List[tuple] = []          # (count, start index in out)
    for raw in lines:
        for ln in str(raw).split("\n"):
            t = ln.split("#", 1)[0].strip()
            if not t:
                continue
            if t.upper().startswith("REPEAT"):
                m = re.match(r"^REPEAT\s+(\d+)\s*\{$", t, re.I)
                if not m:
                    raise RuntimeError(f"cannot parse '{t}'")
                stack.append((int(m.group(1)), len(out)))
            elif t == "}":
                if not stack:
                    raise RuntimeError("unbalanced '}' in circuit")
                count, start = stack.pop()
                body = out[start:]
                for _ in range(count - 1):
                    out.extend(body)
            else:
                out.append(t)
    if stack:
        raise RuntimeError("unterminated REPEAT block")
    return out


def compile_circuit(lines: Iterable[str]) -> CircuitProgram:
    """Stim text (list of instruction strings, as ``StorageSim.circuit``) -> :class:`CircuitProgram`."""
    ops: List[tuple] = []
    thr: List[int] = [0]
    thr_index = {0: 0}
    nq = nmeas = nsites = 0
    detectors: List[List[int]] = []
    observables: dict = {}
    for t in _unroll(list(lines)):
        m = _LINE.match(t)
        if not m:
            raise RuntimeError(f"cannot parse '{t}'")
        name = m.group(1).upper()
        if name in ("DETECTOR", "OBSERVABLE_INCLUDE"):
            recs = [nmeas + int(k) for k in re.findall(r"rec\[(-\d+)\]", m.group(3))]
            if any(r < 0 for r in recs):
                raise RuntimeError(f"measurement record lookback out of range in '{t}'")
            if name == "DETECTOR":
                detectors.append(recs)
            else:
                observables.setdefault(int(float(m.group(2))), []).extend(recs)
            continue
        if name in _ANNOTATIONS:
            continue
        name = _ALIASES.get(name, name)
        if name not in OPCODES:
            raise RuntimeError(f"instruction '{m.group(1)}' is not in the storage-simulation gate set")
        code = OPCODES[name]
        p = float(m.group(2)) if m.group(2) not in (None, "") else 0.0
        if not 0.0 <= p <= 1.0:
            raise RuntimeError(f"probability out of range in '{t}'")
        targets = [int(x) for x in m.group(3).split()] if m.group(3) else []
        if code in (7, 8, 9, 10) and p == 0.0:
            continue                                            # a noise channel of strength 0
        ti = 0
        if p > 0.0:
            v = int(min(np.floor(p * 4294967296.0), 4294967295.0))
            if v not in thr_index:
                thr_index[v] = len(thr)
                thr.append(v)
            ti = thr_index[v]
        two = code in (1, 2, 8)
        if two and len(targets) % 2:
            raise RuntimeError(f"odd number of targets in '{t}'")
        step = 2 if two else 1
        for k in range(0, len(targets), step):
            q0 = targets[k]
            q1 = targets[k + 1] if two else 0
            nq = max(nq, q0 + 1, q1 + 1 if two else 0)
            site = 0
            if ti:
                site = nsites
                nsites += 1
            ops.append((code | (ti << 8), q0, q1, site))
            nmeas += 3 <= code <= 6
    arr = np.asarray(ops, dtype=np.int32).reshape(-1, 4)
    obs = [observables.get(i, []) for i in range(max(observables) + 1)] if observables else []
    return CircuitProgram(np.ascontiguousarray(arr), np.asarray(thr, dtype=np.uint32), int(nq), int(nmeas), int(nsites),
                          detectors, obs)


class CircuitSampler:
    """``stim``-sampler stand-in on the GPU: ``sample(shots)`` -> ``(shots, num_measurements)`` 0/1."""

    def __init__(self, circuit, device: int = 0):
        self.program = circuit if isinstance(circuit, CircuitProgram) else compile_circuit(circuit)
        self.device = int(device)
        self._h = C.c_void_p()
        p = self.program
        _lib.check(_lib.lib().qd_circuit_create(p.ops.shape[0], _lib.ptr(p.ops), p.thresholds.shape[0], _lib.ptr(p.thresholds),
                                                p.num_qubits, p.num_measurements, self.device, C.byref(self._h)))

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qd_circuit_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def num_measurements(self) -> int:
        return self.program.num_measurements

    def sample_device(self, seed, shot0, shots, d_records, stride, stream=None):
        _lib.check(_lib.lib().qd_circuit_sample(self._h, int(seed), int(shot0), int(shots), _lib.ptr(d_records), int(stride), stream))

    def sample(self, shots: int, seed: int = 0, shot0: int = 0, bit_packed: bool = False) -> np.ndarray:
        import torch

        from .decoders import row_words, unpack_bits
        dev = torch.device(f"cuda:{self.device}")
        w = row_words(self.program.num_measurements)
        rec = torch.zeros((shots, w), dtype=torch.int32, device=dev)
        self.sample_device(seed, shot0, shots, rec, w, torch.cuda.current_stream(dev).cuda_stream)
        torch.cuda.synchronize(dev)
        packed = rec.cpu().numpy().view(np.uint32)
        return packed if bit_packed else unpack_bits(packed, self.program.num_measurements).astype(bool)


# ---------------------------------------------------------------------------------------------
# detector error model: what `circuit.detector_error_model()` + DetectorSpacetimeCode give the decoder
# ---------------------------------------------------------------------------------------------
def depolarize1_independent_probability(p: float) -> float:
    """DEPOLARIZE1(p) as three independent X / Y / Z flips of probability q each: q (1 - q) = p / 3."""
    return 0.5 - 0.5 * np.sqrt(1.0 - 4.0 * p / 3.0)


def depolarize2_independent_probability(p: float) -> float:
    """DEPOLARIZE2(p) as fifteen independent two-qubit Pauli flips of equal probability."""
    return 0.5 - 0.5 * (1.0 - 16.0 * p / 15.0) ** 0.125


@dataclass
class DetectorErrorModel:
    """``fault_check_matrix`` (detectors x faults, CSR uint32), ``fault_map`` (observables x faults) and
    ``fault_priors`` — the three fields of the reference's ``DetectorSpacetimeCode``
    (``python/qldpc/spacetime_code.py:122-183``) with row = detector / observable id (the reference's
    ``lil_to_coo`` puts the *position in the list* there, ``:168-171``, which scrambles both matrices)."""
    fault_check_matrix: object
    fault_map: object
    fault_priors: np.ndarray
    detector_matrix: object      # detectors x measurements: which record bits a detector XORs
    observable_matrix: object    # observables x measurements


def detector_error_model(circuit, device: int = 0) -> DetectorErrorModel:
    """Every elementary fault of the noisy circuit (each Pauli of each DEPOLARIZE target, each measurement
    flip, each X/Z_ERROR) is propagated on its own through the noiseless circuit on the GPU
    (``qd_circuit_propagate_faults``); faults with the same symptom (detectors, observables) merge with
    ``p = p1 (1 - p2) + p2 (1 - p1)``; faults without symptom are dropped.  Depolarising channels enter as
    independent Pauli flips (the convention of Stim's error analysis)."""
    import torch
    from scipy import sparse

    from .decoders import row_words, unpack_bits
    prog = circuit if isinstance(circuit, CircuitProgram) else compile_circuit(circuit)
    if not prog.detectors:
        raise ValueError("circuit has no DETECTOR annotations")
    thr_p = prog.thresholds.astype(np.float64) / 4294967296.0
    faults, probs = [], []
    for code_ti, _q0, _q1, site in prog.ops:
        code, ti = int(code_ti) & 0xff, int(code_ti) >> 8
        if not ti:
            continue
        p = float(thr_p[ti])
        if code == 7:
            q = depolarize1_independent_probability(p)
            faults += [(site, c) for c in (1, 2, 3)]
            probs += [q] * 3
        elif code == 8:
            q = depolarize2_independent_probability(p)
            faults += [(site, c) for c in range(1, 16)]
            probs += [q] * 15
        else:
            faults.append((site, 0))
            probs.append(p)
    nmeas, ndet, nobs = prog.num_measurements, len(prog.detectors), len(prog.observables)

    def incidence(rows, n_rows):
        ij = [(i, r) for i, recs in enumerate(rows) for r in recs]
        a = sparse.coo_matrix((np.ones(len(ij), dtype=np.int64), ([i for i, _ in ij], [j for _, j in ij])), shape=(n_rows, nmeas))
        a = a.tocsr()
        a.data %= 2                                                  # a record named twice cancels
        a.eliminate_zeros()
        return a
    dmat, omat = incidence(prog.detectors, ndet), incidence(prog.observables, nobs)
    smp = CircuitSampler(prog, device)
    dev = torch.device(f"cuda:{device}")
    f = torch.tensor(np.asarray(faults, dtype=np.int32).reshape(-1, 2), device=dev)
    w = row_words(nmeas)
    rec = torch.zeros((f.shape[0], w), dtype=torch.int32, device=dev)
    _lib.check(_lib.lib().qd_circuit_propagate_faults(smp._h, f.shape[0], _lib.ptr(f), _lib.ptr(rec), w,
                                                      torch.cuda.current_stream(dev).cuda_stream))
    torch.cuda.synchronize(dev)
    records = sparse.csr_matrix(unpack_bits(rec.cpu().numpy().view(np.uint32), nmeas).astype(np.int64))
    det = (records @ dmat.T).tocsr()
    det.data %= 2
    det.eliminate_zeros()
    det.sort_indices()
    obs = (records @ omat.T).tocsr() if nobs else sparse.csr_matrix((len(faults), 0), dtype=np.int64)
    obs.data %= 2
    obs.eliminate_zeros()
    obs.sort_indices()
    merged: dict = {}
    for k in range(len(faults)):
        key = (tuple(int(v) for v in det.indices[det.indptr[k]:det.indptr[k + 1]]),
               tuple(int(v) for v in obs.indices[obs.indptr[k]:obs.indptr[k + 1]]))
        if not key[0] and not key[1]:
            continue
        pk = probs[k]
        old = merged.get(key)
        merged[key] = pk if old is None else old * (1.0 - pk) + pk * (1.0 - old)
    keys = sorted(merged)                                            # deterministic column order
    fi = [(d, j) for j, (ds, _) in enumerate(keys) for d in ds]
    oi = [(o, j) for j, (_, os_) in enumerate(keys) for o in os_]
    mk = lambda ij, shape: sparse.coo_matrix((np.ones(len(ij), dtype=np.uint32), ([a for a, _ in ij], [b for _, b in ij])),  # noqa: E731
                                             shape=shape, dtype=np.uint32).tocsr()
    return DetectorErrorModel(mk(fi, (ndet, len(keys))), mk(oi, (nobs, len(keys))),
                              np.array([merged[k] for k in keys], dtype=np.float64), dmat.astype(np.uint32), omat.astype(np.uint32))

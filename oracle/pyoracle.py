"""ctypes front-end of the CPU oracle (``oracle/libqldpc_oracle.so``).

TEST INFRASTRUCTURE.  Importable only from ``tests/``, ``__graft_entry__.smoke()`` and
``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product package
``exp_ldpc_b200`` must never import this module (``tests/test_boundary.py`` greps for it).

Parity status: see ``oracle/qldpc_oracle.h`` — structural pieces pinned on reference-generated
fixtures; BP and SSF numerics are "parity unpinned" (``ldpc`` is absent, SSF has no reference).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np
from scipy import sparse

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libqldpc_oracle.so")

METHODS = {"ps": 0, "prod_sum": 0, "product_sum": 0, "ms": 3, "msl": 3, "min_sum": 3, "min_sum_log": 3}


def build(force: bool = False) -> str:
    """Compile the oracle with gcc (``make -C oracle``) if the .so is missing or stale."""
    srcs = [os.path.join(_HERE, f) for f in ("qldpc_oracle.c", "qldpc_oracle_osd.c", "qldpc_oracle.h", "Makefile")]
    stale = (not os.path.exists(_LIB_PATH)) or any(
        os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs if os.path.exists(s))
    if force or stale:
        subprocess.run(["make", "-C", _HERE] + (["-B"] if force else []), check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        i32p, u8p, u32p, f64p, i64p = (C.POINTER(t) for t in (C.c_int32, C.c_uint8, C.c_uint32, C.c_double, C.c_int64))
        L.orc_philox4x32_10.argtypes = [u32p, u32p, u32p]
        L.orc_bernoulli_threshold.argtypes = [C.c_double]
        L.orc_bernoulli_threshold.restype = C.c_uint32
        L.orc_sample_bits.argtypes = [C.c_uint64, C.c_uint32, C.c_int64, C.c_int64, C.c_int, u32p, u8p]
        L.orc_syndrome.argtypes = [C.c_int, i32p, i32p, u8p, u8p]
        L.orc_logical_fail.argtypes = [C.c_int, C.c_int, u8p, u8p]
        L.orc_logical_fail.restype = C.c_int
        L.orc_bp_create.argtypes = [C.c_int, C.c_int, i32p, i32p, f64p, C.c_int, C.c_int, C.c_double]
        L.orc_bp_create.restype = C.c_void_p
        L.orc_bp_destroy.argtypes = [C.c_void_p]
        L.orc_bp_decode.argtypes = [C.c_void_p, u8p, u8p, f64p, C.POINTER(C.c_int)]
        L.orc_bp_decode.restype = C.c_int
        L.orc_bp_soft.argtypes = [C.c_void_p]
        L.orc_bp_soft.restype = f64p
        L.orc_osd0.argtypes = [C.c_int, C.c_int, i32p, i32p, u8p, f64p, u8p]
        L.orc_osd0.restype = C.c_int
        L.orc_osd.argtypes = [C.c_int, C.c_int, i32p, i32p, u8p, f64p, C.c_int, C.c_int, u8p]
        L.orc_osd.restype = C.c_int
        L.orc_bposd_decode.argtypes = [C.c_void_p, u8p, C.c_int, C.c_int, u8p, C.POINTER(C.c_int)]
        L.orc_bposd_decode.restype = C.c_int
        L.orc_bposd0_decode.argtypes = [C.c_void_p, u8p, u8p, C.POINTER(C.c_int)]
        L.orc_bposd0_decode.restype = C.c_int
        L.orc_circuit_sample.argtypes = [C.c_int, i32p, u32p, C.c_int, C.c_int, C.c_uint64, C.c_int64, C.c_int64, u8p]
        L.orc_circuit_propagate_faults.argtypes = [C.c_int, i32p, C.c_int, C.c_int, C.c_int64, i32p, u8p]
        L.orc_ssf_create.argtypes = [C.c_int, C.c_int, i32p, i32p, C.c_int, i32p, i32p]
        L.orc_ssf_create.restype = C.c_void_p
        L.orc_ssf_destroy.argtypes = [C.c_void_p]
        L.orc_ssf_decode.argtypes = [C.c_void_p, u8p, C.c_int, u8p, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_ssf_decode.restype = C.c_int
        L.orc_run_code_capacity.argtypes = [C.c_void_p, C.c_int, C.c_uint64, C.c_int64, C.c_int64, C.c_int,
                                            i64p, u8p, u8p, i32p, u8p]
        L.orc_run_code_capacity.restype = C.c_int
        L.orc_num_threads.restype = C.c_int
        _lib = L
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _csr(h):
    h = sparse.csr_matrix(h)
    h.sum_duplicates()
    h.sort_indices()
    return (h.shape[0], h.shape[1], np.ascontiguousarray(h.indptr, dtype=np.int32),
            np.ascontiguousarray(h.indices, dtype=np.int32))


def philox(ctr, key):
    c = np.asarray(ctr, dtype=np.uint32)
    k = np.asarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_p(c, C.c_uint32), _p(k, C.c_uint32), _p(out, C.c_uint32))
    return out


def thresholds(q, n=None) -> np.ndarray:
    q = np.broadcast_to(np.asarray(q, dtype=np.float64), (n,) if n is not None else np.shape(q))
    return np.array([lib().orc_bernoulli_threshold(float(v)) for v in q], dtype=np.uint32)


def sample_bits(seed: int, stream: int, shot0: int, shots: int, thr: np.ndarray) -> np.ndarray:
    thr = np.ascontiguousarray(thr, dtype=np.uint32)
    out = np.zeros((shots, thr.shape[0]), dtype=np.uint8)
    lib().orc_sample_bits(seed, stream, shot0, shots, thr.shape[0], _p(thr, C.c_uint32), _p(out, C.c_uint8))
    return out


def syndrome(h, vec) -> np.ndarray:
    m, n, ptr, idx = _csr(h)
    v = np.ascontiguousarray(vec, dtype=np.uint8)
    assert v.shape == (n,)
    s = np.zeros(m, dtype=np.uint8)
    lib().orc_syndrome(m, _p(ptr, C.c_int32), _p(idx, C.c_int32), _p(v, C.c_uint8), _p(s, C.c_uint8))
    return s


def logical_fail(L, vec) -> bool:
    L = np.ascontiguousarray(np.asarray(L) & 1, dtype=np.uint8)
    v = np.ascontiguousarray(vec, dtype=np.uint8)
    return bool(lib().orc_logical_fail(L.shape[0], L.shape[1], _p(L, C.c_uint8), _p(v, C.c_uint8)))


class OracleBP:
    """``ldpc.bp_decoder`` restatement (call shape of ``_experiment.py:110-117``)."""

    def __init__(self, h, error_rate=None, channel_probs=None, max_iter=0, bp_method="ps", ms_scaling_factor=0.0):
        self.m, self.n, self._ptr, self._idx = _csr(h)
        if channel_probs is None:
            channel_probs = np.full(self.n, float(error_rate))
        self.probs = np.ascontiguousarray(channel_probs, dtype=np.float64)
        assert self.probs.shape == (self.n,)
        self.method = METHODS[str(bp_method).lower()]
        self._h = lib().orc_bp_create(self.m, self.n, _p(self._ptr, C.c_int32), _p(self._idx, C.c_int32),
                                      _p(self.probs, C.c_double), int(max_iter), self.method, float(ms_scaling_factor))
        self.converge = 0
        self.iter = 0
        self.log_prob_ratios = np.zeros(self.n)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_bp_destroy(self._h)
            self._h = None

    def decode(self, synd) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(synd).astype(np.int64) & 1, dtype=np.uint8)
        assert s.shape == (self.m,)
        dec = np.zeros(self.n, dtype=np.uint8)
        it = C.c_int(0)
        self.converge = lib().orc_bp_decode(self._h, _p(s, C.c_uint8), _p(dec, C.c_uint8),
                                            _p(self.log_prob_ratios, C.c_double), C.byref(it))
        self.iter = it.value
        return dec.astype(np.int64)


class OracleBPOSD0(OracleBP):
    """``ldpc.bposd_decoder(..., osd_method='osd0')`` restatement: BP, then OSD-0 if BP did not converge."""

    def decode(self, synd) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(synd).astype(np.int64) & 1, dtype=np.uint8)
        assert s.shape == (self.m,)
        dec = np.zeros(self.n, dtype=np.uint8)
        it = C.c_int(0)
        self.converge = lib().orc_bposd0_decode(self._h, _p(s, C.c_uint8), _p(dec, C.c_uint8), C.byref(it))
        self.iter = it.value
        self.soft = np.ctypeslib.as_array(lib().orc_bp_soft(self._h), shape=(self.n,)).copy()
        return dec.astype(np.int64)


OSD_METHODS = {"osd0": 0, "osd_0": 0, "osd_e": 1, "osd_cs": 2}


class OracleBPOSD(OracleBP):
    """``ldpc.bposd_decoder(..., osd_method=, osd_order=)`` restatement."""

    def __init__(self, h, osd_method="osd0", osd_order=0, **kw):
        super().__init__(h, **kw)
        self.osd_method, self.osd_order = OSD_METHODS[osd_method], int(osd_order)

    def decode(self, synd) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(synd).astype(np.int64) & 1, dtype=np.uint8)
        dec = np.zeros(self.n, dtype=np.uint8)
        it = C.c_int(0)
        self.converge = lib().orc_bposd_decode(self._h, _p(s, C.c_uint8), self.osd_method, self.osd_order, _p(dec, C.c_uint8),
                                               C.byref(it))
        self.iter = it.value
        return dec.astype(np.int64)


def osd0(h, synd, key_ascending, method="osd0", order=0) -> np.ndarray:
    """OSD solution for an explicit ranking key (ascending = most likely in error first)."""
    m, n, ptr, idx = _csr(h)
    s = np.ascontiguousarray(np.asarray(synd).astype(np.int64) & 1, dtype=np.uint8)
    key = np.ascontiguousarray(key_ascending, dtype=np.float64)
    dec = np.zeros(n, dtype=np.uint8)
    ok = lib().orc_osd(m, n, _p(ptr, C.c_int32), _p(idx, C.c_int32), _p(s, C.c_uint8), _p(key, C.c_double),
                       OSD_METHODS[method], int(order), _p(dec, C.c_uint8))
    if not ok:
        raise ValueError("syndrome is not in the column space of H")
    return dec


def circuit_sample(ops, thresholds, num_qubits, num_measurements, seed, shot0, shots) -> np.ndarray:
    """Pauli-frame sampler on the host: ``(shots, num_measurements)`` 0/1 records."""
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
    thr = np.ascontiguousarray(thresholds, dtype=np.uint32)
    rec = np.zeros((shots, num_measurements), dtype=np.uint8)
    lib().orc_circuit_sample(ops.shape[0], _p(ops, C.c_int32), _p(thr, C.c_uint32), int(num_qubits), int(num_measurements),
                             int(seed), int(shot0), int(shots), _p(rec, C.c_uint8))
    return rec


def circuit_propagate_faults(ops, num_qubits, num_measurements, site_and_pauli) -> np.ndarray:
    """Records of the noiseless circuit with exactly one forced fault each: ``(nfaults, num_measurements)``."""
    ops = np.ascontiguousarray(ops, dtype=np.int32).reshape(-1, 4)
    f = np.ascontiguousarray(site_and_pauli, dtype=np.int32).reshape(-1, 2)
    rec = np.zeros((f.shape[0], num_measurements), dtype=np.uint8)
    lib().orc_circuit_propagate_faults(ops.shape[0], _p(ops, C.c_int32), int(num_qubits), int(num_measurements), f.shape[0],
                                       _p(f, C.c_int32), _p(rec, C.c_uint8))
    return rec


class OracleSSF:
    def __init__(self, hz, hx):
        self.m, self.n, self._zp, self._zi = _csr(hz)
        self.mg, n2, self._xp, self._xi = _csr(hx)
        assert n2 == self.n
        self._h = lib().orc_ssf_create(self.m, self.n, _p(self._zp, C.c_int32), _p(self._zi, C.c_int32),
                                       self.mg, _p(self._xp, C.c_int32), _p(self._xi, C.c_int32))
        if not self._h:
            raise ValueError("SSF oracle: generator weight > 12 or check window > 64")
        self.converge = 0
        self.rounds = 0
        self.final_weight = 0

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_ssf_destroy(self._h)
            self._h = None

    def decode(self, synd, max_rounds: int = 0) -> np.ndarray:
        s = np.ascontiguousarray(np.asarray(synd).astype(np.int64) & 1, dtype=np.uint8)
        dec = np.zeros(self.n, dtype=np.uint8)
        r, w = C.c_int(0), C.c_int(0)
        self.converge = lib().orc_ssf_decode(self._h, _p(s, C.c_uint8), int(max_rounds), _p(dec, C.c_uint8),
                                             C.byref(r), C.byref(w))
        self.rounds, self.final_weight = r.value, w.value
        return dec.astype(np.int64)


class _Experiment(C.Structure):
    _fields_ = [("m", C.c_int), ("n", C.c_int), ("k", C.c_int),
                ("row_ptr", C.POINTER(C.c_int32)), ("col_idx", C.POINTER(C.c_int32)),
                ("Lz", C.POINTER(C.c_uint8)), ("thr", C.POINTER(C.c_uint32)), ("probs", C.POINTER(C.c_double)),
                ("max_iter", C.c_int), ("method", C.c_int), ("alpha", C.c_double),
                ("mg", C.c_int), ("hx_row_ptr", C.POINTER(C.c_int32)), ("hx_col_idx", C.POINTER(C.c_int32)),
                ("osd_order", C.c_int)]


def run_code_capacity(hz, lz, q_sample, prior, *, decoder="bp", hx=None, max_iter=0, bp_method="ms",
                      ms_scaling_factor=0.0, seed=0, shot0=0, shots=1, threads=0, per_shot=False, osd_order=0):
    """Sample -> syndrome -> decode -> logical flag for ``shots`` shots on the host cores.
    Returns ``dict(failures, shots, non_converged, iters_total[, decodings, converged, iters, fail])``."""
    m, n, ptr, idx = _csr(hz)
    lz = np.ascontiguousarray(np.asarray(lz) & 1, dtype=np.uint8).reshape(-1, n)
    thr = thresholds(q_sample, n)
    probs = np.ascontiguousarray(np.broadcast_to(np.asarray(prior, dtype=np.float64), (n,)))
    x = _Experiment()
    x.m, x.n, x.k = m, n, lz.shape[0]
    x.row_ptr, x.col_idx = _p(ptr, C.c_int32), _p(idx, C.c_int32)
    x.Lz, x.thr, x.probs = _p(lz, C.c_uint8), _p(thr, C.c_uint32), _p(probs, C.c_double)
    x.max_iter, x.method, x.alpha = int(max_iter), METHODS.get(str(bp_method).lower(), 3), float(ms_scaling_factor)
    x.osd_order = int(osd_order)
    keep = [ptr, idx, lz, thr, probs]
    if decoder == "ssf":
        mg, _, xp, xi = _csr(hx)
        x.mg, x.hx_row_ptr, x.hx_col_idx = mg, _p(xp, C.c_int32), _p(xi, C.c_int32)
        keep += [xp, xi]
    counters = np.zeros(4, dtype=np.int64)
    dec = conv = its = fail = None
    if per_shot:
        dec = np.zeros((shots, n), dtype=np.uint8)
        conv = np.zeros(shots, dtype=np.uint8)
        its = np.zeros(shots, dtype=np.int32)
        fail = np.zeros(shots, dtype=np.uint8)
    null8, null32 = C.POINTER(C.c_uint8)(), C.POINTER(C.c_int32)()
    rc = lib().orc_run_code_capacity(
        C.byref(x), {"bp": 0, "ssf": 1, "bposd0": 2, "bposd_e": 3, "bposd_cs": 4}[decoder], int(seed), int(shot0), int(shots), int(threads),
        _p(counters, C.c_int64),
        _p(dec, C.c_uint8) if per_shot else null8, _p(conv, C.c_uint8) if per_shot else null8,
        _p(its, C.c_int32) if per_shot else null32, _p(fail, C.c_uint8) if per_shot else null8)
    if rc != 0:
        raise RuntimeError("oracle experiment failed (unsupported SSF geometry?)")
    out = dict(failures=int(counters[0]), shots=int(counters[1]), non_converged=int(counters[2]),
               iters_total=int(counters[3]))
    if per_shot:
        out.update(decodings=dec, converged=conv, iters=its, fail=fail)
    return out


def num_threads() -> int:
    return int(lib().orc_num_threads())

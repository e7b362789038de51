"""Decoder objects with the call shape the reference uses — backed by the sm_100a kernels.

The reference builds ``ldpc.bp_decoder(H, channel_probs=..., max_iter=..., bp_method=...,
ms_scaling_factor=...)`` once per worker and calls ``.decode(syndrome)`` once per shot
(``python/qldpc/misc/_experiment.py:110-117,137-149``).  :class:`bp_decoder` keeps that exact
surface (ctor keywords, ``decode``, ``converge`` / ``iter`` / ``log_prob_ratios`` side outputs) and
adds the batched entry points that replace the per-shot Python loop
(``_experiment.py:200-209``): ``decode_batch`` (host arrays) and ``decode_batch_device`` (device
pointers, optional fused logical-failure epilogue).

:class:`ssf_decoder` is the small-set-flip decoder of the north star; the reference has no
counterpart (SURVEY.md Appendix C is its specification).

All compute goes through ``libqldpc_b200.so``; there is no CPU path in this module.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import numpy as np

from . import _lib
from .codes import canonical_csr

__all__ = [
    "pack_bits", "unpack_bits", "row_words", "TannerGraph", "Logicals", "Sampler", "bp_decoder", "bposd_decoder", "ssf_decoder",
    "bernoulli_thresholds",
]

_BP_METHODS = {
    # ldpc.bp_decoder's accepted spellings ("ms" is routed to the log-domain min-sum there)
    "ps": 0, "prod_sum": 0, "product_sum": 0, "prod sum": 0, "0": 0,
    "ms": 3, "min_sum": 3, "minimum_sum": 3, "min sum": 3, "minimum sum": 3, "1": 3,
    "msl": 3, "min_sum_log": 3, "minimum_sum_log": 3, "ms_log": 3, "3": 3,
}


def row_words(nbits: int, align: int = 4) -> int:
    """Words per packed row, padded to ``align`` words (16-byte rows for 128-bit loads)."""
    w = (nbits + 31) // 32
    return (w + align - 1) // align * align


def pack_bits(bits, stride: Optional[int] = None) -> np.ndarray:
    """``(shots, nbits)`` 0/1 array -> ``(shots, stride)`` uint32, bit j -> word j>>5, bit j&31."""
    b = np.ascontiguousarray(np.asarray(bits).astype(np.uint8) & 1)
    if b.ndim == 1:
        b = b[None, :]
    shots, nbits = b.shape
    stride = row_words(nbits) if stride is None else int(stride)
    if stride * 32 < nbits:
        raise ValueError("stride too small")
    by = np.zeros((shots, stride * 4), dtype=np.uint8)
    pk = np.packbits(b, axis=1, bitorder="little")
    by[:, :pk.shape[1]] = pk
    return by.view("<u4").astype(np.uint32, copy=False)


def unpack_bits(words, nbits: int) -> np.ndarray:
    w = np.ascontiguousarray(np.asarray(words, dtype=np.uint32))
    if w.ndim == 1:
        w = w[None, :]
    by = w.astype("<u4", copy=False).view(np.uint8)
    return np.unpackbits(by, axis=1, bitorder="little")[:, :nbits]


def bernoulli_thresholds(q, n: Optional[int] = None) -> np.ndarray:
    """uint32 thresholds of the Philox sampler: bit = (u32 < floor(q * 2^32)), clamped."""
    q = np.broadcast_to(np.asarray(q, dtype=np.float64), (n,) if n is not None else np.shape(q))
    t = np.floor(np.where(q > 0, q, 0.0) * 4294967296.0)
    return np.minimum(t, 4294967295.0).astype(np.uint32)


class _Handle:
    _destroy = None

    def __init__(self):
        self._h = C.c_void_p()

    @property
    def handle(self):
        return self._h

    def close(self):
        if getattr(self, "_h", None) and self._h.value and self._destroy:
            getattr(_lib.lib(), self._destroy)(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TannerGraph(_Handle):
    """Device-resident Tanner graph of a parity-check matrix (CSR + CSC edge orders)."""
    _destroy = "qd_graph_destroy"

    def __init__(self, h, device: int = 0):
        super().__init__()
        h = canonical_csr(h)
        self.m, self.n = h.shape
        self.nnz = int(h.nnz)
        self.device = int(device)
        self.indptr = np.ascontiguousarray(h.indptr, dtype=np.int32)
        self.indices = np.ascontiguousarray(h.indices, dtype=np.int32)
        _lib.check(_lib.lib().qd_graph_create(self.m, self.n, self.nnz, _lib.ptr(self.indptr), _lib.ptr(self.indices),
                                              self.device, C.byref(self._h)))

    def syndrome_batch_device(self, d_vec, vec_stride, shots, d_synd, synd_stride, stream=None):
        _lib.check(_lib.lib().qd_syndrome_batch(self._h, _lib.ptr(d_vec), vec_stride, shots, _lib.ptr(d_synd),
                                                synd_stride, stream))


class Logicals(_Handle):
    _destroy = "qd_logicals_destroy"

    def __init__(self, dense, device: int = 0):
        super().__init__()
        L = np.ascontiguousarray((np.asarray(dense) & 1).astype(np.uint8))
        if L.ndim != 2:
            raise ValueError("logicals must be k x n")
        self.k, self.n = L.shape
        _lib.check(_lib.lib().qd_logicals_create(self.k, self.n, _lib.ptr(L), int(device), C.byref(self._h)))

    def fail_batch_device(self, d_a, d_b, stride, shots, d_fail, d_count=None, stream=None):
        _lib.check(_lib.lib().qd_logical_fail_batch(self._h, _lib.ptr(d_a), _lib.ptr(d_b), stride, shots,
                                                    _lib.ptr(d_fail), _lib.ptr(d_count), stream))


class Sampler(_Handle):
    """Philox4x32-10 Bernoulli sampler keyed by (seed, global shot index): rank-count independent."""
    _destroy = "qd_sampler_destroy"

    def __init__(self, flip_probability, n: Optional[int] = None, device: int = 0):
        super().__init__()
        self.thresholds = np.ascontiguousarray(bernoulli_thresholds(flip_probability, n))
        self.nbits = int(self.thresholds.shape[0])
        _lib.check(_lib.lib().qd_sampler_create(self.nbits, _lib.ptr(self.thresholds), int(device), C.byref(self._h)))

    def run_device(self, seed, stream_id, shot0, shots, d_bits, stride, stream=None):
        _lib.check(_lib.lib().qd_sampler_run(self._h, int(seed), int(stream_id), int(shot0), int(shots),
                                             _lib.ptr(d_bits), int(stride), stream))


def _as_syndrome_batch(x, m: int, packed: bool, nwm: int):
    if packed:
        s = np.ascontiguousarray(np.asarray(x, dtype=np.uint32))
        if s.ndim != 2 or s.shape[1] < nwm:
            raise ValueError("packed syndromes must be (shots, >= ceil(m/32)) uint32")
        return s
    a = np.asarray(x)
    if a.ndim == 1:
        a = a[None, :]
    if a.ndim != 2 or a.shape[1] != m:
        raise ValueError(f"syndromes must have shape (shots, {m})")
    return pack_bits(a.astype(np.int64) & 1)


class bp_decoder:
    """Drop-in for ``ldpc.bp_decoder`` (syndrome decoding) running on the GPU.

    Parameters mirror upstream: ``parity_check_matrix``, ``error_rate`` *or* ``channel_probs``,
    ``max_iter`` (0 -> n), ``bp_method`` ('ps' | 'ms' | 'msl' and upstream's aliases),
    ``ms_scaling_factor`` (0 -> 1 - 2**-iteration).  Unknown keywords are ignored, as upstream
    does (the reference passes OSD options to ``bp_decoder`` too, ``_experiment.py:110-113``).

    Extra keywords: ``precision`` ('f64' reproduces the reference's double-precision operation
    order bit for bit; 'f32' is faster, not reproducible), ``device``, ``kernel`` ('auto' |
    'regular' | 'regular_cluster' | 'generic' | 'cluster' | 'global': which CUDA kernel family decodes; the results are identical).
    """

    def __init__(self, parity_check_matrix, error_rate=None, max_iter=0, bp_method=0, ms_scaling_factor=1.0,
                 channel_probs=(None,), precision="f64", device=0, kernel="auto", **_ignored):
        self.graph = parity_check_matrix if isinstance(parity_check_matrix, TannerGraph) else TannerGraph(
            parity_check_matrix, device)
        self.m, self.n = self.graph.m, self.graph.n
        cp = None
        if channel_probs is not None:
            arr = np.asarray(channel_probs, dtype=object).ravel()
            if arr.size and arr[0] is not None:
                cp = np.asarray(channel_probs, dtype=np.float64).ravel()
                if cp.shape[0] != self.n:
                    raise ValueError(
                        f"The length of the channel probability vector must be eqaul to the block length n={self.n}.")
        if cp is None:
            if error_rate is None:
                raise ValueError("Please specify the error channel. Either: 1) error_rate: float or 2) channel_probs: "
                                 "list of floats indicating the error probability on each bit.")
            cp = np.full(self.n, float(error_rate), dtype=np.float64)
        self.channel_probs = np.ascontiguousarray(cp)
        self.error_rate = error_rate
        key = str(bp_method).lower()
        if key not in _BP_METHODS:
            raise ValueError(f"BP method '{bp_method}' is invalid. Please choose from the following methods: "
                             "'product_sum', 'minimum_sum', 'product_sum_log' or 'minimum_sum_log'")
        self.bp_method = _BP_METHODS[key]
        self.max_iter = int(max_iter) if int(max_iter) > 0 else self.n
        self.ms_scaling_factor = float(ms_scaling_factor)
        self.precision = {"f64": 64, "f32": 32, 64: 64, 32: 32}[precision]
        self._h = C.c_void_p()
        _lib.check(_lib.lib().qd_bp_create(self.graph.handle, _lib.ptr(self.channel_probs), int(max_iter), self.bp_method,
                                           self.ms_scaling_factor, self.precision, C.byref(self._h)))
        if kernel != "auto":
            _lib.check(_lib.lib().qd_bp_select_kernel(self._h, {"global": 5, "generic": 6, "regular": 7, "cluster": 8, "regular_cluster": 9}[kernel]))
        self.converge = 0
        self.iter = 0
        self.bp_decoding = np.zeros(self.n, dtype=np.int64)
        self.log_prob_ratios = np.zeros(self.n, dtype=np.float64)

    # ---- lifetime
    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qd_bp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    # ---- launch geometry (tuning)
    def configure(self, threads=0, slots=0, ctas_per_sm=0, teams=0):
        _lib.check(_lib.lib().qd_bp_configure(self._h, threads, slots, ctas_per_sm, teams))
        return self.config()

    def config(self) -> dict:
        v = [C.c_int32() for _ in range(6)]
        _lib.check(_lib.lib().qd_bp_get_config(self._h, *[C.byref(x) for x in v]))
        out = dict(zip(("threads", "slots", "ctas_per_sm", "teams", "smem_bytes", "grid"), (x.value for x in v)))
        k = [C.c_int32() for _ in range(3)]
        _lib.check(_lib.lib().qd_bp_get_kernel(self._h, *[C.byref(x) for x in k]))
        out["kernel"] = {5: "global", 6: "generic", 7: "regular", 8: "cluster", 9: "regular_cluster"}[k[0].value]
        if out["kernel"] == "regular_cluster":      # in this family `ctas_per_sm` carries the cluster size (CTAs per group of shots)
            out["cluster_ctas"] = out["ctas_per_sm"]
        return out

    # ---- the ldpc call
    def decode(self, input_vector) -> np.ndarray:
        s = np.asarray(input_vector)
        if s.shape != (self.m,):
            raise ValueError(f"The input to the ldpc.bp_decoder.decode must be a syndrome (of length={self.m}). "
                             f"The inputted vector has length={s.shape[0] if s.ndim else 0}. Valid formats are `np.ndarray`.")
        sb = np.ascontiguousarray(s.astype(np.int64) & 1, dtype=np.uint8)
        out = np.zeros(self.n, dtype=np.uint8)
        cv, it = C.c_int32(0), C.c_int32(0)
        _lib.check(_lib.lib().qd_bp_decode_one(self._h, _lib.ptr(sb), _lib.ptr(out), C.byref(cv), C.byref(it),
                                               _lib.ptr(self.log_prob_ratios)))
        self.converge, self.iter = int(cv.value), int(it.value)
        self.bp_decoding = out.astype(np.int64)
        return self.bp_decoding

    # ---- batched replacements of the per-shot loop
    def decode_batch(self, syndromes, packed: bool = False, return_packed: bool = False):
        """Host arrays in, host arrays out: ``(corrections, converged, iters)``.
        ``syndromes``: ``(shots, m)`` 0/1 (or packed uint32 rows when ``packed``)."""
        nwm, nwn = (self.m + 31) // 32, (self.n + 31) // 32
        s = _as_syndrome_batch(syndromes, self.m, packed, nwm)
        shots = s.shape[0]
        cstride = row_words(self.n)
        corr = np.zeros((shots, cstride), dtype=np.uint32)
        conv = np.zeros(shots, dtype=np.uint8)
        its = np.zeros(shots, dtype=np.int32)
        _lib.check(_lib.lib().qd_bp_decode_batch_host(self._h, _lib.ptr(s), s.shape[1], shots, _lib.ptr(corr), cstride,
                                                      _lib.ptr(conv), _lib.ptr(its)))
        if return_packed:
            return corr, conv, its
        return unpack_bits(corr, self.n), conv, its

    def decode_batch_device(self, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged=None, d_iters=None,
                            d_llr=None, d_err=None, err_stride=0, logicals: Optional[Logicals] = None, d_fail=None,
                            d_counters=None, stream=None):
        """Device pointers (torch CUDA tensors or raw ints).  Asynchronous on ``stream``."""
        _lib.check(_lib.lib().qd_bp_decode_batch(
            self._h, _lib.ptr(d_synd), synd_stride, shots, _lib.ptr(d_corr), corr_stride, _lib.ptr(d_converged),
            _lib.ptr(d_iters), _lib.ptr(d_llr), _lib.ptr(d_err), err_stride,
            logicals.handle if logicals is not None else None, _lib.ptr(d_fail), _lib.ptr(d_counters), stream))


class bposd_decoder(bp_decoder):
    """Drop-in for ``ldpc.bposd_decoder`` (``_experiment.py:23-27,96-100``): belief propagation, then
    ordered-statistics post-processing on the shots BP did not converge on.

    ``osd_method`` 'osd0' (information set = the first rank(H) independent columns in the order of the
    BP soft output), 'osd_e' (exhaustive over the first ``osd_order`` other columns) and 'osd_cs'
    (combination sweep, the reference's default with ``osd_order = 7``, ``_experiment.py:218-219``).
    Side outputs as upstream: ``bp_decoding`` (BP's own hard decision), ``osdw_decoding`` = the returned
    vector, ``converge`` / ``iter`` describe the BP stage."""

    def __init__(self, parity_check_matrix, error_rate=None, max_iter=0, bp_method=0, ms_scaling_factor=1.0,
                 channel_probs=(None,), osd_order=-1, osd_method="osd0", precision="f64", device=0, kernel="auto",
                 **_ignored):
        key = str(osd_method).lower()
        if key in ("osd_e", "osde", "exhaustive", "e"):
            method = 1
        elif key in ("osd_cs", "1", "cs", "combination_sweep"):
            method = 2
        elif key in ("osd0", "osd_0", "0", "zero"):
            method = 0
        else:
            raise ValueError(f"ERROR: OSD method '{osd_method}' invalid. Please choose from the following methods: "
                             "'OSD_0', 'OSD_E' or 'OSD_CS'.")
        order = max(int(osd_order), 0) if method else 0
        super().__init__(parity_check_matrix, error_rate=error_rate, max_iter=max_iter, bp_method=bp_method,
                         ms_scaling_factor=ms_scaling_factor, channel_probs=channel_probs, precision=precision,
                         device=device, kernel=kernel)
        self.osd_method = ("osd0", "osd_e", "osd_cs")[method if order else 0]
        self.osd_order = order
        self._osd = C.c_void_p()
        _lib.check(_lib.lib().qd_osd_create(self.graph.handle, method, order, C.byref(self._osd)))
        self.osd0_decoding = np.zeros(self.n, dtype=np.int64)
        self.osdw_decoding = np.zeros(self.n, dtype=np.int64)

    def close(self):
        if getattr(self, "_osd", None) and self._osd.value:
            _lib.lib().qd_osd_destroy(self._osd)
            self._osd = C.c_void_p()
        super().close()

    @property
    def osd_handle(self):
        return self._osd

    @property
    def rank(self) -> int:
        return int(_lib.lib().qd_osd_rank(self._osd))

    def decode_batch(self, syndromes, packed: bool = False, return_packed: bool = False):
        nwm = (self.m + 31) // 32
        s = _as_syndrome_batch(syndromes, self.m, packed, nwm)
        shots = s.shape[0]
        cstride = row_words(self.n)
        corr = np.zeros((shots, cstride), dtype=np.uint32)
        conv = np.zeros(shots, dtype=np.uint8)
        its = np.zeros(shots, dtype=np.int32)
        _lib.check(_lib.lib().qd_bposd_decode_batch_host(self._h, self._osd, _lib.ptr(s), s.shape[1], shots, _lib.ptr(corr),
                                                         cstride, _lib.ptr(conv), _lib.ptr(its)))
        if return_packed:
            return corr, conv, its
        return unpack_bits(corr, self.n), conv, its

    def decode(self, input_vector) -> np.ndarray:
        s = np.asarray(input_vector)
        if s.shape != (self.m,):
            raise ValueError(f"The input to the ldpc.bp_decoder.decode must be a syndrome (of length={self.m}). "
                             f"The inputted vector has length={s.shape[0] if s.ndim else 0}. Valid formats are `np.ndarray`.")
        corr, conv, its = self.decode_batch(s[None, :])
        self.converge, self.iter = int(conv[0]), int(its[0])
        self.osd0_decoding = self.osdw_decoding = corr[0].astype(np.int64)
        return self.osdw_decoding

    def decode_batch_device(self, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged=None, d_iters=None,
                            d_llr=None, d_err=None, err_stride=0, logicals: Optional[Logicals] = None, d_fail=None,
                            d_counters=None, stream=None):
        if d_llr is not None:
            raise ValueError("bposd_decoder.decode_batch_device does not export soft outputs")
        _lib.check(_lib.lib().qd_bposd_decode_batch(
            self._h, self._osd, _lib.ptr(d_synd), synd_stride, shots, _lib.ptr(d_corr), corr_stride, _lib.ptr(d_converged),
            _lib.ptr(d_iters), _lib.ptr(d_err), err_stride, logicals.handle if logicals is not None else None,
            _lib.ptr(d_fail), _lib.ptr(d_counters), stream))


class ssf_decoder:
    """Small-set-flip decoder: ``ssf_decoder(hz, hx).decode(syndrome)``; candidate sets are the
    subsets of the supports of the rows of ``hx``."""

    def __init__(self, hz, hx, max_rounds: int = 0, device: int = 0):
        self.graph = hz if isinstance(hz, TannerGraph) else TannerGraph(hz, device)
        self.m, self.n = self.graph.m, self.graph.n
        gx = canonical_csr(hx)
        if gx.shape[1] != self.n:
            raise ValueError("x and z checks act on an inconsistent number of qubits")
        self._gp = np.ascontiguousarray(gx.indptr, dtype=np.int32)
        self._gi = np.ascontiguousarray(gx.indices, dtype=np.int32)
        self.max_rounds = int(max_rounds)
        self._h = C.c_void_p()
        _lib.check(_lib.lib().qd_ssf_create(self.graph.handle, gx.shape[0], _lib.ptr(self._gp), _lib.ptr(self._gi),
                                            self.max_rounds, C.byref(self._h)))
        self.converge = 0
        self.iter = 0

    def close(self):
        if getattr(self, "_h", None) and self._h.value:
            _lib.lib().qd_ssf_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def handle(self):
        return self._h

    def decode_batch(self, syndromes, packed: bool = False, return_packed: bool = False):
        nwm = (self.m + 31) // 32
        s = _as_syndrome_batch(syndromes, self.m, packed, nwm)
        shots = s.shape[0]
        cstride = row_words(self.n)
        corr = np.zeros((shots, cstride), dtype=np.uint32)
        conv = np.zeros(shots, dtype=np.uint8)
        rounds = np.zeros(shots, dtype=np.int32)
        _lib.check(_lib.lib().qd_ssf_decode_batch_host(self._h, _lib.ptr(s), s.shape[1], shots, _lib.ptr(corr), cstride,
                                                       _lib.ptr(conv), _lib.ptr(rounds)))
        if return_packed:
            return corr, conv, rounds
        return unpack_bits(corr, self.n), conv, rounds

    def decode(self, syndrome) -> np.ndarray:
        s = np.asarray(syndrome)
        if s.shape != (self.m,):
            raise ValueError(f"syndrome must have length {self.m}")
        corr, conv, rounds = self.decode_batch(s[None, :])
        self.converge, self.iter = int(conv[0]), int(rounds[0])
        return corr[0].astype(np.int64)

    def decode_batch_device(self, d_synd, synd_stride, shots, d_corr, corr_stride, d_converged=None, d_rounds=None,
                            d_err=None, err_stride=0, logicals: Optional[Logicals] = None, d_fail=None,
                            d_counters=None, stream=None):
        _lib.check(_lib.lib().qd_ssf_decode_batch(
            self._h, _lib.ptr(d_synd), synd_stride, shots, _lib.ptr(d_corr), corr_stride, _lib.ptr(d_converged),
            _lib.ptr(d_rounds), _lib.ptr(d_err), err_stride, logicals.handle if logicals is not None else None,
            _lib.ptr(d_fail), _lib.ptr(d_counters), stream))

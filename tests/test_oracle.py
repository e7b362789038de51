"""The CPU oracle against known answers and an independent pure-Python restatement (small cases).
CPU only.  BP/SSF numerics are 'parity unpinned' w.r.t. the real `ldpc` package (absent here);
what is pinned is listed in oracle/qldpc_oracle.h."""
import math
import os

import numpy as np
import pytest

from conftest import GOLDEN
from oracle import pyoracle as po


def test_philox_known_answers():
    # Random123 kat_vectors, philox4x32-10
    assert [hex(v) for v in po.philox([0, 0, 0, 0], [0, 0])] == ['0x6627e8d5', '0xe169c58d', '0xbc57ac4c', '0x9b00dbd8']
    assert [hex(v) for v in po.philox([0xffffffff] * 4, [0xffffffff] * 2)] == ['0x408f276d', '0x41c83b0e', '0xa20bc7c6', '0x6d5451fd']
    assert [hex(v) for v in po.philox([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])] == \
        ['0xd16cfe09', '0x94fdcceb', '0x5001e420', '0x24126ea1']


def test_sampler_is_counter_based():
    thr = po.thresholds(0.1, 37)
    a = po.sample_bits(5, 0, 0, 64, thr)
    b = po.sample_bits(5, 0, 32, 32, thr)
    assert np.array_equal(a[32:], b)                       # independent of where the range starts
    assert not np.array_equal(a, po.sample_bits(6, 0, 0, 64, thr))
    big = po.sample_bits(1, 0, 0, 20000, po.thresholds(0.05, 64))
    assert abs(big.mean() - 0.05) < 0.002
    assert po.thresholds([0.0, 1.0, 0.5]).tolist() == [0, 4294967295, 2147483648]


def test_syndrome_and_logical_flag_match_numpy(code225):
    rng = np.random.default_rng(0)
    hz, lz = code225.checks.z, code225.logicals.z
    for _ in range(20):
        e = rng.integers(0, 2, hz.shape[1]).astype(np.uint8)
        assert np.array_equal(po.syndrome(hz, e), (hz @ e.astype(np.int64)) % 2)     # _experiment.py:124
        assert po.logical_fail(lz, e) == bool(((lz.astype(np.int64) @ e) % 2).any())  # _experiment.py:209


# ---- independent pure-Python restatement of the two BP variants -------
def _py_bp(h, probs, synd, max_iter, method, alpha0):
    h = h.tocsr()
    m, n = h.shape
    rows = [list(h.indices[h.indptr[i]:h.indptr[i + 1]]) for i in range(m)]
    cols = [[] for _ in range(n)]
    for i, r in enumerate(rows):
        for j in r:
            cols[j].append(i)
    b2c, c2b = {}, {}
    for j in range(n):
        for i in cols[j]:
            b2c[i, j] = probs[j] / (1 - probs[j]) if method == "ps" else math.log((1 - probs[j]) / probs[j])
    dec = np.zeros(n, dtype=np.uint8)
    for it in range(1, max_iter + 1):
        alpha = alpha0 if alpha0 != 0 else 1.0 - 2.0 ** (-it)
        for i in range(m):
            r = rows[i]
            if method == "ps":
                t = -1.0 if synd[i] else 1.0
                for j in r:
                    c2b[i, j] = t
                    t *= 2.0 / (1.0 + b2c[i, j]) - 1.0
                t = 1.0
                for j in reversed(r):
                    x = c2b[i, j] * t
                    c2b[i, j] = (1.0 - x) / (1.0 + x)
                    t *= 2.0 / (1.0 + b2c[i, j]) - 1.0
            else:
                t, sg, sgn = 1e308, int(synd[i]), {}
                for j in r:
                    c2b[i, j] = t
                    sgn[j] = sg
                    t = min(t, abs(b2c[i, j]))
                    sg += b2c[i, j] <= 0
                t, sg = 1e308, 0
                for j in reversed(r):
                    c2b[i, j] = min(c2b[i, j], t)
                    sgn[j] += sg
                    c2b[i, j] *= (-1.0 if sgn[j] & 1 else 1.0) * alpha
                    t = min(t, abs(b2c[i, j]))
                    sg += b2c[i, j] <= 0
        for j in range(n):
            c = cols[j]
            if method == "ps":
                t = probs[j] / (1 - probs[j])
                for i in c:
                    b2c[i, j] = t
                    t *= c2b[i, j]
                    t = 1.0 if math.isnan(t) else t
                dec[j] = t >= 1
                t = 1.0
                for i in reversed(c):
                    b2c[i, j] *= t
                    t *= c2b[i, j]
                    t = 1.0 if math.isnan(t) else t
            else:
                t = math.log((1 - probs[j]) / probs[j])
                for i in c:
                    b2c[i, j] = t
                    t += c2b[i, j]
                dec[j] = t <= 0
                t = 0.0
                for i in reversed(c):
                    b2c[i, j] += t
                    t += c2b[i, j]
        if np.array_equal((h @ dec.astype(np.int64)) % 2, synd):
            return dec, 1, it
    return dec, 0, max_iter


@pytest.mark.parametrize("method,alpha", [("ms", 0.0), ("ms", 0.625), ("ms", 1.0), ("ps", 0.0)])
def test_bp_matches_python_restatement_bit_for_bit(code225, method, alpha):
    hz = code225.checks.z
    n = hz.shape[1]
    p = 0.04
    dec = po.OracleBP(hz, error_rate=p, max_iter=12, bp_method=method, ms_scaling_factor=alpha)
    errs = po.sample_bits(3, 0, 0, 12, po.thresholds(0.05, n))
    for e in errs:
        s = po.syndrome(hz, e)
        got = dec.decode(s)
        want, conv, it = _py_bp(hz, np.full(n, p), s, 12, method, alpha)
        assert np.array_equal(got, want) and dec.converge == conv and dec.iter == it


@pytest.mark.parametrize("method", ["ms", "ps"])
def test_bp_known_answers(code225, method):
    hz = code225.checks.z
    m, n = hz.shape
    dec = po.OracleBP(hz, error_rate=0.01, max_iter=0, bp_method=method, ms_scaling_factor=0.0)
    # zero syndrome -> zero correction after one iteration
    out = dec.decode(np.zeros(m, dtype=np.uint8))
    assert not out.any() and dec.converge == 1 and dec.iter == 1
    # every weight-1 error is recovered exactly
    for j in range(0, n, 7):
        e = np.zeros(n, dtype=np.uint8)
        e[j] = 1
        out = dec.decode(po.syndrome(hz, e))
        assert dec.converge == 1 and np.array_equal(out, e), j
    # max_iter = 0 means n iterations (ldpc convention; _experiment.py:224 passes n anyway)
    hard = po.sample_bits(11, 0, 0, 40, po.thresholds(0.12, n))
    its = []
    for e in hard:
        dec.decode(po.syndrome(hz, e))
        its.append(dec.iter if not dec.converge else 0)
    assert max(its) == n


def test_bp_channel_probs_vector(code225):
    hz = code225.checks.z
    n = hz.shape[1]
    probs = np.linspace(0.01, 0.08, n)
    d = po.OracleBP(hz, channel_probs=probs, max_iter=20, bp_method="ms")
    e = po.sample_bits(2, 0, 0, 1, po.thresholds(0.03, n))[0]
    s = po.syndrome(hz, e)
    out = d.decode(s)
    want, conv, it = _py_bp(hz, probs, s, 20, "ms", 0.0)
    assert np.array_equal(out, want) and d.iter == it
    assert np.allclose(d.log_prob_ratios[out == 0] > 0, True)


# ---- SSF ---------------------------------------------------------------------------------------
def _py_ssf(hz, hx, synd):
    hz = hz.tocsc()
    hx = hx.tocsr()
    sigma = synd.copy().astype(np.int64)
    n = hz.shape[1]
    dec = np.zeros(n, dtype=np.uint8)
    rounds = 0
    while sigma.any():
        best = None
        for g in range(hx.shape[0]):
            supp = hx.indices[hx.indptr[g]:hx.indptr[g + 1]]
            for mask in range(1, 1 << len(supp)):
                f = [supp[i] for i in range(len(supp)) if (mask >> i) & 1]
                tog = np.zeros_like(sigma)
                for q in f:
                    tog[hz.indices[hz.indptr[q]:hz.indptr[q + 1]]] ^= 1
                delta = int(sigma.sum() - (sigma ^ tog).sum())
                if delta > 0 and (best is None or delta * best[1] > best[0] * len(f)):
                    best = (delta, len(f), f, tog)
        if best is None:
            break
        for q in best[2]:
            dec[q] ^= 1
        sigma ^= best[3]
        rounds += 1
    return dec, int(not sigma.any()), rounds


def test_ssf_matches_python_restatement(code225):
    hz, hx = code225.checks.z, code225.checks.x
    d = po.OracleSSF(hz, hx)
    errs = po.sample_bits(9, 0, 0, 3, po.thresholds(0.02, hz.shape[1]))
    for e in errs:
        s = po.syndrome(hz, e)
        got = d.decode(s)
        want, conv, rounds = _py_ssf(hz, hx, s)
        assert np.array_equal(got, want) and d.converge == conv and d.rounds == rounds


def test_ssf_properties(code225):
    hz, hx = code225.checks.z, code225.checks.x
    d = po.OracleSSF(hz, hx)
    assert not d.decode(np.zeros(hz.shape[0], dtype=np.uint8)).any() and d.converge == 1 and d.rounds == 0
    errs = po.sample_bits(4, 0, 0, 200, po.thresholds(0.02, hz.shape[1]))
    for e in errs:
        s = po.syndrome(hz, e)
        out = d.decode(s)
        resid = po.syndrome(hz, out.astype(np.uint8)) ^ s
        assert int(resid.sum()) == d.final_weight and (d.converge == 1) == (d.final_weight == 0)
        assert d.rounds <= int(s.sum())                     # every flip lowers the weight by >= 1
    capped = d.decode(po.syndrome(hz, errs[0]), max_rounds=1)
    assert capped.sum() <= 7 and d.rounds <= 1


def test_batched_experiment_equals_single_shot_calls(code225):
    hz, hx, lz = code225.checks.z, code225.checks.x, code225.logicals.z
    n = hz.shape[1]
    q = 0.03
    for mode in ("bp", "ssf"):
        r = po.run_code_capacity(hz, lz, q, q, decoder=mode, hx=hx, max_iter=10 if mode == "bp" else 0, bp_method="ms",
                                 seed=5, shot0=100, shots=50, threads=3, per_shot=True)
        errs = po.sample_bits(5, 0, 100, 50, po.thresholds(q, n))
        dec = po.OracleBP(hz, error_rate=q, max_iter=10, bp_method="ms") if mode == "bp" else po.OracleSSF(hz, hx)
        fails = 0
        for s in range(50):
            out = dec.decode(po.syndrome(hz, errs[s]))
            assert np.array_equal(out, r["decodings"][s])
            f = po.logical_fail(lz, errs[s] ^ out.astype(np.uint8))
            assert f == bool(r["fail"][s])
            fails += f
        assert fails == r["failures"] and r["shots"] == 50
        # thread count must not change anything
        r1 = po.run_code_capacity(hz, lz, q, q, decoder=mode, hx=hx, max_iter=10 if mode == "bp" else 0, bp_method="ms",
                                  seed=5, shot0=100, shots=50, threads=1, per_shot=True)
        assert np.array_equal(r1["decodings"], r["decodings"]) and r1["failures"] == r["failures"]


def test_committed_decoder_vectors(code225, code1600):
    """Regression vectors of the oracle itself (tests/golden/decoder_vectors.npz, written by
    tests/golden/make_decoder_vectors.py); the GPU tests compare against the same file."""
    path = os.path.join(GOLDEN, "decoder_vectors.npz")
    g = np.load(path)
    from tests_golden_cases import CASES, run_case
    for name in CASES:
        out = run_case(name, {"hgp225": code225, "hgp1600": code1600})
        assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"]), name
        assert np.array_equal(out["iters"], g[name + "_iters"]) and np.array_equal(out["fail"], g[name + "_fail"])


# ---------------------------------------------------------------------------------------------
# OSD-0 (ldpc.bposd_decoder, osd_method "osd0"): the oracle against an independent numpy elimination
# ---------------------------------------------------------------------------------------------
def _numpy_osd0(h, synd, key):
    order = np.lexsort((np.arange(h.shape[1]), key))           # ascending key, ties by column index
    a = (h.toarray()[:, order] % 2).astype(np.uint8)
    b = np.asarray(synd, dtype=np.uint8).copy()
    m = a.shape[0]
    used = np.zeros(m, bool)
    pcol = -np.ones(m, int)
    for c in range(a.shape[1]):
        rows = np.where((a[:, c] == 1) & ~used)[0]
        if len(rows) == 0:
            continue
        r = rows[0]
        used[r], pcol[r] = True, c
        for i in np.where(a[:, c] == 1)[0]:
            if i != r:
                a[i] ^= a[r]
                b[i] ^= b[r]
    x = np.zeros(h.shape[1], np.uint8)
    for r in range(m):
        if pcol[r] >= 0:
            x[order[pcol[r]]] = b[r]
    return x


def test_osd0_matches_numpy_elimination(code225):
    from oracle import pyoracle as po
    hz = code225.checks.z
    rng = np.random.default_rng(0)
    for trial in range(12):
        e = (rng.random(hz.shape[1]) < 0.06).astype(np.uint8)
        s = (hz @ e) % 2
        key = rng.normal(size=hz.shape[1]) if trial % 3 else np.round(rng.normal(size=hz.shape[1]), 0)   # with ties
        x = po.osd0(hz, s, key)
        assert np.array_equal((hz @ x) % 2, s)
        assert np.array_equal(x, _numpy_osd0(hz, s, key))
    # zero syndrome -> zero; information set prefers the low keys
    assert not po.osd0(hz, np.zeros(hz.shape[0], np.uint8), rng.normal(size=hz.shape[1])).any()


def test_bposd0_keeps_converged_bp_and_fixes_the_rest(code225):
    from oracle import pyoracle as po
    hz = code225.checks.z
    rng = np.random.default_rng(5)
    bp = po.OracleBP(hz, error_rate=0.03, max_iter=12, bp_method="ms", ms_scaling_factor=0.0)
    bo = po.OracleBPOSD0(hz, error_rate=0.03, max_iter=12, bp_method="ms", ms_scaling_factor=0.0)
    seen = set()
    for _ in range(60):
        e = (rng.random(hz.shape[1]) < 0.05).astype(np.uint8)
        s = (hz @ e) % 2
        a, b = bp.decode(s), bo.decode(s)
        assert bo.converge == bp.converge and bo.iter == bp.iter
        assert np.array_equal((hz @ b) % 2, s)                      # BP+OSD always returns a valid correction
        if bp.converge:
            assert np.array_equal(a, b)
        else:
            assert np.array_equal(b, po.osd0(hz, s, bo.soft))
        seen.add(bp.converge)
    assert seen == {0, 1}



def test_higher_order_osd_never_heavier_and_valid(code225):
    from oracle import pyoracle as po
    hz = code225.checks.z
    rng = np.random.default_rng(3)
    for _ in range(6):
        e = (rng.random(hz.shape[1]) < 0.07).astype(np.uint8)
        s = (hz @ e) % 2
        key = rng.normal(size=hz.shape[1])
        w0 = po.osd0(hz, s, key).sum()
        for method, order in (("osd_e", 6), ("osd_cs", 7)):
            x = po.osd0(hz, s, key, method, order)
            assert np.array_equal((hz @ x) % 2, s) and x.sum() <= w0
        assert np.array_equal(po.osd0(hz, s, key, "osd_cs", 0), po.osd0(hz, s, key))

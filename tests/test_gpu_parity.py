"""Parity tests proper: the CUDA path, called through the C ABI, against the CPU oracle on identical
seeded inputs and against the committed golden vectors.  Bar: bit-exact for syndromes, logical
flags, SSF and the fp64 BP path (hard decisions, iteration counts, convergence flags);
>= 99.99 % is the north-star floor for BP — the fp64 path is held to 100 %."""
import os

import numpy as np
import pytest

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(built):
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def codes(code225, code1600):
    return {"hgp225": code225, "hgp1600": code1600}


def test_sampler_bit_exact(gpu, code225):
    import torch
    from exp_ldpc_b200.decoders import Sampler, row_words, unpack_bits
    from oracle import pyoracle as po
    for n, q in ((225, 0.05), (37, 0.3), (1600, 0.01)):
        stride = row_words(n)
        smp = Sampler(q, n)
        for shot0, shots, seed, stream in ((0, 300, 0, 0), (2 ** 33 + 5, 64, 12345678901234567, 3)):
            d = torch.full((shots, stride), 0xFFFFFFFF - 2 ** 32, dtype=torch.int32, device=gpu)
            smp.run_device(seed, stream, shot0, shots, d, stride)
            torch.cuda.synchronize()
            got = d.cpu().numpy().view(np.uint32)
            want = po.sample_bits(seed, stream, shot0, shots, po.thresholds(q, n))
            assert np.array_equal(unpack_bits(got, n), want)
            assert not (got[:, (n + 31) // 32:] != 0).any()                       # padding words are zero
            if n % 32:
                assert not (got[:, n // 32] >> (n % 32)).any()                    # padding bits are zero


@pytest.mark.parametrize("which", ["hgp225", "hgp1600"])
def test_syndrome_and_logical_flag_bit_exact(gpu, codes, which):
    import torch
    from exp_ldpc_b200.decoders import Logicals, TannerGraph, pack_bits, row_words, unpack_bits
    from oracle import pyoracle as po
    code = codes[which]
    hz, lz = code.checks.z, code.logicals.z
    m, n = hz.shape
    shots = 257
    rng = np.random.default_rng(1)
    e = rng.integers(0, 2, size=(shots, n)).astype(np.uint8)
    e[0] = 0
    e[1] = 1
    c = rng.integers(0, 2, size=(shots, n)).astype(np.uint8)
    g = TannerGraph(hz)
    L = Logicals(lz)
    es, ss = row_words(n), row_words(m)
    d_e = torch.from_numpy(pack_bits(e, es).view(np.int32)).to(gpu)
    d_c = torch.from_numpy(pack_bits(c, es).view(np.int32)).to(gpu)
    d_s = torch.empty((shots, ss), dtype=torch.int32, device=gpu)
    d_f = torch.empty(shots, dtype=torch.uint8, device=gpu)
    d_n = torch.zeros(1, dtype=torch.int64, device=gpu)
    g.syndrome_batch_device(d_e, es, shots, d_s, ss)
    L.fail_batch_device(d_e, d_c, es, shots, d_f, d_n)
    torch.cuda.synchronize()
    got = unpack_bits(d_s.cpu().numpy().view(np.uint32), m)
    want = np.stack([po.syndrome(hz, e[s]) for s in range(shots)])
    assert np.array_equal(got, want)
    wantf = np.array([po.logical_fail(lz, e[s] ^ c[s]) for s in range(shots)], dtype=np.uint8)
    assert np.array_equal(d_f.cpu().numpy(), wantf) and int(d_n.item()) == int(wantf.sum())
    # linearity: syndrome(a ^ b) = syndrome(a) ^ syndrome(b)
    d_x = torch.bitwise_xor(d_e, d_c)
    d_s2 = torch.empty_like(d_s)
    d_s3 = torch.empty_like(d_s)
    g.syndrome_batch_device(d_c, es, shots, d_s2, ss)
    g.syndrome_batch_device(d_x, es, shots, d_s3, ss)
    torch.cuda.synchronize()
    assert torch.equal(d_s3, torch.bitwise_xor(d_s, d_s2))


def _run_gpu_case(name, codes, precision="f64", config=None, kernel="auto"):
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from tests_golden_cases import CASES, case_probabilities
    key, decoder, opts, p, shots, seed = CASES[name]
    q, prior = case_probabilities(p)
    exp = CodeCapacityExperiment(codes[key], q, prior, decoder, opts, precision=precision, kernel=kernel)
    if decoder == "bp" and kernel != "auto":
        assert exp.decoder.config()["kernel"] == kernel
    if config and decoder == "bp":
        from exp_ldpc_b200._lib import QdError, QD_EUNSUP
        try:
            exp.decoder.configure(**config)
        except ValueError as e:          # geometry not representable for this graph
            pytest.skip(str(e))
        except QdError as e:             # (cluster mode: too few rows for that many CTAs, or no room)
            if e.code != QD_EUNSUP:
                raise
            pytest.skip(str(e))
    return exp.run(shots, seed=seed, per_shot=True)


def _all_case_names():
    from tests_golden_cases import CASES
    return list(CASES)


def _bp_case_names():
    from tests_golden_cases import CASES
    return [k for k, v in CASES.items() if v[1] == "bp"]


@pytest.mark.parametrize("name", _bp_case_names())
def test_generic_bp_kernel_matches_committed_golden_vectors(gpu, codes, name):
    """The HGP codes are row-regular and take the row-regular kernel by default; the generic kernel
    (irregular graphs) must give the same bits."""
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, kernel="generic")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])
    assert np.array_equal(out["converged"], g[name + "_conv"])


def test_hgp_codes_take_the_row_regular_kernel(gpu, codes):
    from exp_ldpc_b200.decoders import bp_decoder
    for key in ("hgp225", "hgp1600"):
        for prec in ("f64", "f32"):
            for method in ("ms", "ps"):
                cfg = bp_decoder(codes[key].checks.z, error_rate=0.02, max_iter=8, bp_method=method, precision=prec).config()
                assert cfg["kernel"] == "regular", (key, prec, method, cfg)


@pytest.mark.parametrize("name", _all_case_names())
def test_decoders_match_committed_golden_vectors(gpu, codes, name):
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes)
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])
    assert np.array_equal(out["converged"], g[name + "_conv"])
    assert out["failures"] == int(g[name + "_fail"].sum()) and out["shots"] == len(g[name + "_fail"])
    assert out["non_converged"] == int((g[name + "_conv"] == 0).sum()) and out["iters_total"] == int(g[name + "_iters"].sum())


@pytest.mark.parametrize("name", ["ms_sched_225", "ps_225", "ms_sched_1600"])
@pytest.mark.parametrize("config", [dict(threads=32, slots=1, ctas_per_sm=1), dict(threads=96, slots=3, ctas_per_sm=2),
                                    dict(threads=1024, slots=1, ctas_per_sm=1), dict(threads=256, slots=3, ctas_per_sm=1),
                                    dict(threads=512, slots=4, ctas_per_sm=1, teams=2), dict(threads=384, slots=3, ctas_per_sm=1, teams=3),
                                    dict(threads=1024, slots=4, ctas_per_sm=1, teams=4)])
def test_bp_result_independent_of_launch_geometry(gpu, codes, name, config):
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, config=config, kernel="generic")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])


@pytest.mark.parametrize("name", ["ms_sched_225", "ps_225", "ms_sched_1600"])
@pytest.mark.parametrize("config", [dict(threads=32, slots=1, ctas_per_sm=1), dict(threads=64, slots=4, ctas_per_sm=2),
                                    dict(threads=128, slots=8, ctas_per_sm=1), dict(threads=800, slots=2, ctas_per_sm=1),
                                    dict(threads=256, slots=4, ctas_per_sm=1), dict(threads=288, slots=1, ctas_per_sm=2)])
def test_regular_bp_kernel_independent_of_launch_geometry(gpu, codes, name, config):
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, config=config, kernel="regular")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])


@pytest.mark.parametrize("method,alpha,p", [("ms", 0.0, 0.01), ("ms", 0.0, 0.04), ("ms", 0.8, 0.02), ("ps", 0.0, 0.03)])
def test_bp_fresh_seeds_against_live_oracle(gpu, code225, method, alpha, p):
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 5000
    opts = dict(max_iter=40, bp_method=method, ms_scaling_factor=alpha)
    got = CodeCapacityExperiment(code225, q, prior, "bp", opts).run(shots, seed=99, shot0=10 ** 6, per_shot=True)
    ref = po.run_code_capacity(code225.checks.z, code225.logicals.z, q, prior, decoder="bp", max_iter=40, bp_method=method,
                               ms_scaling_factor=alpha, seed=99, shot0=10 ** 6, shots=shots, per_shot=True)
    agree = (got["decodings"] == ref["decodings"]).all(axis=1).mean()
    assert agree == 1.0, f"hard decisions agree on {agree:.6%} of shots"
    assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])
    assert got["failures"] == ref["failures"]


def test_bp_decoder_duck_type_matches_oracle(gpu, code225):
    """The ldpc.bp_decoder call shape used at _experiment.py:110-117."""
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    hz = code225.checks.z
    m, n = hz.shape
    probs = np.linspace(0.01, 0.05, n)
    opts = dict(max_iter=n, bp_method="ps", ms_scaling_factor=0, osd_method="osd_cs", osd_order=7)   # OSD keys ignored
    dec = bp_decoder(hz.tocoo(), channel_probs=probs, **opts)
    ref = po.OracleBP(hz, channel_probs=probs, max_iter=n, bp_method="ps", ms_scaling_factor=0)
    errs = po.sample_bits(1, 0, 0, 20, po.thresholds(0.03, n))
    for e in errs:
        s = po.syndrome(hz, e).astype(np.float64)            # the reference hands in float64 0/1
        out = dec.decode(s)
        want = ref.decode(s)
        assert out.shape == (n,) and out.dtype.kind == "i" and np.array_equal(out, want)
        assert dec.converge == ref.converge and dec.iter == ref.iter
        assert np.allclose(dec.log_prob_ratios, ref.log_prob_ratios, rtol=1e-12, atol=1e-12)
    zero = dec.decode(np.zeros(m))
    assert not zero.any() and dec.converge == 1 and dec.iter == 1
    with pytest.raises(ValueError):
        dec.decode(np.zeros(m + 1))
    with pytest.raises(ValueError):
        bp_decoder(hz, channel_probs=np.full(n + 1, 0.1))
    with pytest.raises(ValueError):
        bp_decoder(hz)
    with pytest.raises(ValueError):
        bp_decoder(hz, error_rate=0.1, bp_method="nope")


def test_decode_batch_host_api(gpu, code225):
    from exp_ldpc_b200.decoders import bp_decoder, ssf_decoder
    from oracle import pyoracle as po
    hz, hx = code225.checks.z, code225.checks.x
    m, n = hz.shape
    errs = po.sample_bits(8, 0, 0, 3000, po.thresholds(0.03, n))
    synd = np.stack([po.syndrome(hz, e) for e in errs])
    dec = bp_decoder(hz, error_rate=0.02, max_iter=25, bp_method="ms", ms_scaling_factor=0.0)
    ref = po.OracleBP(hz, error_rate=0.02, max_iter=25, bp_method="ms", ms_scaling_factor=0.0)
    corr, conv, its = dec.decode_batch(synd)
    assert corr.shape == (3000, n)
    for s in range(0, 3000, 17):
        want = ref.decode(synd[s])
        assert np.array_equal(corr[s], want) and conv[s] == ref.converge and its[s] == ref.iter
    # converged shots reproduce the syndrome (H c = s)
    ok = conv.astype(bool)
    assert np.array_equal((hz @ corr[ok].T.astype(np.int64)).T % 2, synd[ok])
    # empty and single-shot batches
    c0, v0, i0 = dec.decode_batch(np.zeros((0, m), dtype=np.uint8))
    assert c0.shape == (0, n) and v0.shape == (0,)
    c1, v1, i1 = dec.decode_batch(synd[5])
    assert np.array_equal(c1[0], corr[5])
    # SSF through the same host API
    sd = ssf_decoder(hz, hx)
    rs = po.OracleSSF(hz, hx)
    scorr, sconv, srounds = sd.decode_batch(synd[:400])
    for s in range(400):
        want = rs.decode(synd[s])
        assert np.array_equal(scorr[s], want) and sconv[s] == rs.converge and srounds[s] == rs.rounds


def test_run_simulation_signature_and_counts(gpu, code225):
    from exp_ldpc_b200.experiment import depolarizing_noise, run_simulation
    from oracle import pyoracle as po
    p = 0.02
    prior = lambda p_, *_: 2 / 3 * p_                                         # scripts/p_sweep.py:4-5
    from functools import partial
    opts = {"max_iter": 30, "bp_method": "ms", "ms_scaling_factor": 0.0, "osd_method": "osd_cs", "osd_order": 7}
    flags = run_simulation(1000, code225, partial(prior, p), partial(prior, p), depolarizing_noise, {"p": p, "pm": p},
                           opts, 0, "bp", seed=3)
    assert isinstance(flags, list) and len(flags) == 1000 and isinstance(flags[0], bool)
    q = 2 * p / 3 + p - 4 / 3 * p * p
    ref = po.run_code_capacity(code225.checks.z, code225.logicals.z, q, 2 * p / 3, decoder="bp", max_iter=30,
                               bp_method="ms", seed=3, shots=1000, per_shot=True)
    assert flags == [bool(v) for v in ref["fail"]]
    with pytest.raises(RuntimeError):
        run_simulation(10, code225, partial(prior, p), partial(prior, p), depolarizing_noise, {"p": p, "pm": p}, opts, 0, "nope")
    # the reference's default mode and OSD settings (bposd, osd_cs order 7) run as they are
    flags = run_simulation(400, code225, partial(prior, p), partial(prior, p), depolarizing_noise, {"p": p, "pm": p}, opts, 0,
                           "bposd", seed=3)
    ref = po.run_code_capacity(code225.checks.z, code225.logicals.z, q, 2 * p / 3, decoder="bposd_cs", osd_order=7, max_iter=30,
                               bp_method="ms", seed=3, shots=400, per_shot=True)
    assert flags == [bool(v) for v in ref["fail"]]


def test_full_size_properties_n6400(gpu, code6400):
    """BASELINE config 2 at full code size, size-independent properties: converged shots satisfy
    H c = s; counters add up; results do not depend on how the batch is split."""
    import torch
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    p = 0.01
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    exp = CodeCapacityExperiment(code6400, q, prior, "bp", dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0))
    shots = 3000
    a = exp.run(shots, seed=1, per_shot=True)
    hz = code6400.checks.z
    ok = a["converged"].astype(bool)
    assert ok.sum() > 0 and a["shots"] == shots and a["non_converged"] == int((~ok).sum())
    from oracle import pyoracle as po
    errs = po.sample_bits(1, 0, 0, 64, po.thresholds(q, hz.shape[1]))
    for s in range(64):
        synd = po.syndrome(hz, errs[s])
        if ok[s]:
            assert np.array_equal(po.syndrome(hz, a["decodings"][s].astype(np.uint8)), synd)
    b1 = exp.run(1000, seed=1, shot0=0, per_shot=True)
    b2 = exp.run(2000, seed=1, shot0=1000, per_shot=True)
    assert np.array_equal(np.concatenate([b1["decodings"], b2["decodings"]]), a["decodings"])
    assert b1["failures"] + b2["failures"] == a["failures"] and a["iters_total"] == int(a["iters"].sum())
    # spot-check against the oracle on the first shots
    ref = po.run_code_capacity(hz, code6400.logicals.z, q, prior, decoder="bp", max_iter=32, bp_method="ms", seed=1,
                               shots=48, per_shot=True)
    assert np.array_equal(ref["decodings"], a["decodings"][:48]) and np.array_equal(ref["iters"], a["iters"][:48])


# ---------------------------------------------------------------------------------------------
# the other code families of BASELINE.json (configs 3-4): row degrees 6 / 8 / 9, column tile 4 / 8
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def family_codes():
    from exp_ldpc_b200 import code_builders as cb
    return {name: getattr(cb, name)() for name in ("pk19_b1_882", "pk20_qc_1054", "pk19_b3_1270", "a5_lift_2700")}


@pytest.mark.parametrize("name", ["pk19_b1_882", "pk20_qc_1054", "pk19_b3_1270", "a5_lift_2700"])
@pytest.mark.parametrize("method,alpha,p", [("ms", 0.0, 0.03), ("ms", 0.75, 0.06), ("ps", 0.0, 0.04)])
def test_bp_code_families_against_live_oracle(gpu, family_codes, name, method, alpha, p):
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = family_codes[name]
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 384
    opts = dict(max_iter=24, bp_method=method, ms_scaling_factor=alpha)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bp", max_iter=24, bp_method=method,
                               ms_scaling_factor=alpha, seed=11, shots=shots, per_shot=True)
    for kernel in ("regular", "generic"):
        exp = CodeCapacityExperiment(code, q, prior, "bp", opts, kernel=kernel)
        assert exp.decoder.config()["kernel"] == kernel
        got = exp.run(shots, seed=11, per_shot=True)
        assert (got["decodings"] == ref["decodings"]).all(), (kernel, "hard decisions differ")
        assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])
        assert np.array_equal(got["converged"], ref["converged"])


@pytest.mark.parametrize("name,p", [("pk19_b1_882", 0.01), ("pk20_qc_1054", 0.01), ("a5_lift_2700", 0.006)])
def test_ssf_code_families_against_live_oracle(gpu, family_codes, name, p):
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = family_codes[name]
    q = prior = 2 * p / 3
    shots = 192
    got = CodeCapacityExperiment(code, q, prior, "ssf", {}).run(shots, seed=3, per_shot=True)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="ssf", hx=code.checks.x, seed=3,
                               shots=shots, per_shot=True)
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])
    assert np.array_equal(got["converged"], ref["converged"])


def test_cayley_tanner_lp_against_live_oracle(gpu):
    """BASELINE configs[3]: the [[4320, 28]] Cayley-graph Tanner lifted product over PSL(2,4) (check degrees 5 / 7,
    qubit degrees 2..4; cayley_codes.psl2_tanner_lp_4320) — BP (min-sum, product-sum) and small-set-flip."""
    from exp_ldpc_b200.cayley_codes import psl2_tanner_lp_4320
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = psl2_tanner_lp_4320()
    hz, lz, hx = code.checks.z, code.logicals.z, code.checks.x
    shots = 256
    for method, alpha, p in (("ms", 0.0, 0.03), ("ms", 0.75, 0.06), ("ps", 0.0, 0.04)):
        q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
        opts = dict(max_iter=24, bp_method=method, ms_scaling_factor=alpha)
        ref = po.run_code_capacity(hz, lz, q, prior, decoder="bp", max_iter=24, bp_method=method, ms_scaling_factor=alpha,
                                   seed=11, shots=shots, per_shot=True)
        got = CodeCapacityExperiment(code, q, prior, "bp", opts).run(shots, seed=11, per_shot=True)
        assert (got["decodings"] == ref["decodings"]).all(), (method, alpha, "hard decisions differ")
        assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])
        assert np.array_equal(got["converged"], ref["converged"])
    p = 0.008
    q = prior = 2 * p / 3
    got = CodeCapacityExperiment(code, q, prior, "ssf", {}).run(shots, seed=3, per_shot=True)
    ref = po.run_code_capacity(hz, lz, q, prior, decoder="ssf", hx=hx, seed=3, shots=shots, per_shot=True)
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])
    assert np.array_equal(got["converged"], ref["converged"])
    assert 0 < int(ref["converged"].sum()) and int(ref["iters"].max()) > 1


def test_bp_per_bit_priors_match_oracle(gpu, code225):
    """channel_probs vector (the hybrid decoder's data / measurement priors, _experiment.py:106-113)."""
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    rng = np.random.default_rng(1)
    hz = code225.checks.z
    probs = rng.uniform(0.005, 0.2, hz.shape[1])
    err = (rng.random((200, hz.shape[1])) < probs).astype(np.uint8)
    synd = (err @ hz.T.toarray()) % 2
    for method in ("ms", "ps"):
        ref = po.OracleBP(hz, channel_probs=probs, max_iter=30, bp_method=method, ms_scaling_factor=0.0)
        for kernel in ("regular", "generic"):
            dec = bp_decoder(hz, channel_probs=probs, max_iter=30, bp_method=method, ms_scaling_factor=0.0, kernel=kernel)
            corr, conv, its = dec.decode_batch(synd)
            for s in range(synd.shape[0]):
                want = ref.decode(synd[s])
                assert np.array_equal(corr[s], want), (method, kernel, s)
                assert its[s] == ref.iter and conv[s] == ref.converge


# ---------------------------------------------------------------------------------------------
# multi-round (spacetime) decoding, phenomenological faults: irregular check matrices
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("rounds", [0, 1, 3])
@pytest.mark.parametrize("method", ["ms", "ps"])
def test_spacetime_experiment_against_live_oracle(gpu, code225, rounds, method):
    """H_st = SpacetimeCode(Hz, T) (spacetime_code.py:46-75); faults drawn from the per-column priors;
    failure = Lz . final_correction(f ^ f_hat) (spacetime_code.py:81-84, _experiment.py:208-209)."""
    from exp_ldpc_b200.experiment import SpacetimeExperiment
    from oracle import pyoracle as po
    opts = dict(max_iter=30, bp_method=method, ms_scaling_factor=0.0)
    shots = 256
    data_prior, meas_prior = 0.012, 0.02
    for kernel in ("auto", "generic") + (("regular_cluster",) if rounds else ()):
        exp = SpacetimeExperiment(code225, rounds, data_prior, meas_prior, opts, kernel=kernel)
        if kernel == "auto":
            assert exp.decoder.config()["kernel"] == "regular"      # padded-row variant for T = 3
        if kernel == "regular_cluster":
            assert exp.decoder.config()["ctas_per_sm"] == 2          # two rounds per CTA for T = 3: only measurement columns are cut
        got = exp.run(shots, seed=21, per_shot=True)
        h_st = exp.spacetime.spacetime_check_matrix
        n, m = code225.num_qubits, code225.checks.z.shape[0]
        assert h_st.shape == ((rounds + 1) * m, (rounds + 1) * n + rounds * m)
        lz = np.asarray(code225.logicals.z) & 1
        l_st = np.concatenate([lz] * (rounds + 1) + [np.zeros((lz.shape[0], rounds * m), dtype=lz.dtype)], axis=1)
        ref = po.run_code_capacity(h_st, l_st, exp.prior, exp.prior, decoder="bp", seed=21, shots=shots, per_shot=True, **opts)
        assert (got["decodings"] == ref["decodings"]).all(), (kernel, rounds, method)
        assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["converged"], ref["converged"])
        assert np.array_equal(got["fail"], ref["fail"])
        # the fused failure flag equals the reference recipe: fold both fault vectors, then Lz
        err = po.sample_bits(21, 0, 0, shots, po.thresholds(exp.prior))
        folded = exp.final_corrections(err ^ got["decodings"])
        want = ((folded.astype(np.int64) @ lz.T.astype(np.int64)) % 2).any(axis=1)
        assert np.array_equal(want.astype(np.uint8), got["fail"])


def test_single_shot_matrix_decodes_like_oracle(gpu, code225):
    """[H | I] of SpacetimeCodeSingleShot (spacetime_code.py:20-25): rows of weight 8, columns 1/3/4."""
    from exp_ldpc_b200.decoders import bp_decoder
    from exp_ldpc_b200.spacetime import SpacetimeCodeSingleShot
    from oracle import pyoracle as po
    h = SpacetimeCodeSingleShot(code225.checks.z).spacetime_check_matrix
    rng = np.random.default_rng(3)
    probs = np.concatenate([np.full(code225.num_qubits, 0.02), np.full(h.shape[0], 0.03)])
    err = (rng.random((128, h.shape[1])) < probs).astype(np.uint8)
    synd = (err @ h.T.toarray()) % 2
    ref = po.OracleBP(h, channel_probs=probs, max_iter=20, bp_method="ms", ms_scaling_factor=0.0)
    for kernel in ("auto", "generic"):
        dec = bp_decoder(h, channel_probs=probs, max_iter=20, bp_method="ms", ms_scaling_factor=0.0, kernel=kernel)
        corr, conv, its = dec.decode_batch(synd)
        for s in range(synd.shape[0]):
            assert np.array_equal(corr[s], ref.decode(synd[s])), (kernel, s)
            assert its[s] == ref.iter and conv[s] == ref.converge


# ---------------------------------------------------------------------------------------------
# BP + OSD-0 (ldpc.bposd_decoder with osd_method "osd0"; _experiment.py:23-27,77,82,96-100,125)
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("which,method,p,kernel", [("hgp225", "ms", 0.04, "auto"), ("hgp225", "ps", 0.03, "auto"),
                                                  ("hgp225", "ms", 0.05, "generic"), ("hgp1600", "ms", 0.025, "auto")])
def test_bposd0_experiment_against_live_oracle(gpu, codes, which, method, p, kernel):
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = codes[which]
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 400 if which == "hgp225" else 96
    opts = dict(max_iter=16, bp_method=method, ms_scaling_factor=0.0, osd_method="osd0")
    got = CodeCapacityExperiment(code, q, prior, "bposd", opts, kernel=kernel).run(shots, seed=17, per_shot=True)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bposd0", max_iter=16, bp_method=method,
                               ms_scaling_factor=0.0, seed=17, shots=shots, per_shot=True)
    assert 0 < ref["non_converged"] < shots                        # both branches exercised
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["converged"], ref["converged"]) and np.array_equal(got["iters"], ref["iters"])
    assert np.array_equal(got["fail"], ref["fail"])
    assert got["failures"] == ref["failures"] == int(ref["fail"].sum())
    assert got["non_converged"] == ref["non_converged"]
    # BP + OSD always returns a correction with the right syndrome
    hz = code.checks.z
    err = po.sample_bits(17, 0, 0, shots, po.thresholds(q, hz.shape[1]))
    assert not (((err ^ got["decodings"]) @ hz.T.toarray()) % 2).any()


@pytest.mark.parametrize("which,osd,order", [("hgp1600", "osd0", 0), ("hgp225", "osd_cs", 5)])
def test_osd_two_rows_per_thread_variant(gpu, codes, which, osd, order, monkeypatch):
    """The OSD kernel instantiation that gives a thread two rows (chosen by itself only beyond 1024 checks)
    must produce the same decodings; QD_OSD_RPT forces it on codes the oracle finishes quickly."""
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = codes[which]
    p = 0.03
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 64 if which == "hgp1600" else 300
    opts = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.0, osd_method=osd, osd_order=order)
    monkeypatch.setenv("QD_OSD_RPT", "2")
    got = CodeCapacityExperiment(code, q, prior, "bposd", opts).run(shots, seed=23, per_shot=True)
    monkeypatch.delenv("QD_OSD_RPT")
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bposd0" if order == 0 else "bposd_cs",
                               osd_order=order, max_iter=12, bp_method="ms", ms_scaling_factor=0.0, seed=23, shots=shots,
                               per_shot=True)
    assert ref["non_converged"] > 0
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["fail"], ref["fail"]) and got["failures"] == ref["failures"]


@pytest.mark.parametrize("which,env,osd,order", [("hgp1600", {"QD_OSD_GLOBAL": "1"}, "osd0", 0),
                                                  ("hgp225", {"QD_OSD_GLOBAL": "1", "QD_OSD_RPT": "4"}, "osd_cs", 6),
                                                  ("hgp1600", {"QD_OSD_RPT": "4"}, "osd0", 0)])
def test_osd_global_scratch_and_four_rows_per_thread(gpu, codes, which, env, osd, order, monkeypatch):
    """Codes whose [H | s] exceeds shared memory (n = 6400: 2.5 MB) keep it in a global scratch buffer and give a
    thread up to four rows; both paths are forced here on codes the oracle finishes quickly."""
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = codes[which]
    p = 0.03
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 48 if which == "hgp1600" else 300
    opts = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.0, osd_method=osd, osd_order=order)
    for k, v in env.items():
        monkeypatch.setenv(k, v)
    got = CodeCapacityExperiment(code, q, prior, "bposd", opts).run(shots, seed=29, per_shot=True)
    for k in env:
        monkeypatch.delenv(k)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bposd0" if order == 0 else "bposd_cs",
                               osd_order=order, max_iter=12, bp_method="ms", ms_scaling_factor=0.0, seed=29, shots=shots,
                               per_shot=True)
    assert ref["non_converged"] > 0
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["fail"], ref["fail"]) and got["failures"] == ref["failures"]


@pytest.mark.parametrize("osd,order", [("osd0", 0), ("osd_cs", 5)])
def test_bposd_on_a_code_beyond_shared_memory(gpu, osd, order):
    """A5-lift n = 2700 (m = 1080): [H | s] takes 397 KB; BP + OSD must still give valid, oracle-identical corrections."""
    from exp_ldpc_b200 import code_builders
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = code_builders.a5_lift_2700()
    p = 0.04
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 16
    opts = dict(max_iter=8, bp_method="ms", ms_scaling_factor=0.0, osd_method=osd, osd_order=order)
    got = CodeCapacityExperiment(code, q, prior, "bposd", opts).run(shots, seed=31, per_shot=True)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bposd0" if order == 0 else "bposd_cs",
                               osd_order=order, max_iter=8, bp_method="ms", ms_scaling_factor=0.0, seed=31, shots=shots,
                               per_shot=True)
    assert ref["non_converged"] > 0
    assert (got["decodings"] == ref["decodings"]).all()
    hz = code.checks.z
    err = po.sample_bits(31, 0, 0, shots, po.thresholds(q, hz.shape[1]))
    assert not (((err ^ got["decodings"]) @ hz.T.toarray()) % 2).any()


def test_bposd_decoder_duck_type(gpu, code225):
    """ldpc.bposd_decoder call shape of _experiment.py:96-100,125."""
    from exp_ldpc_b200.decoders import bposd_decoder
    from oracle import pyoracle as po
    hz = code225.checks.z
    dec = bposd_decoder(hz, error_rate=0.03, max_iter=10, bp_method="ms", ms_scaling_factor=0, osd_method="osd0", osd_order=0)
    ref = po.OracleBPOSD0(hz, error_rate=0.03, max_iter=10, bp_method="ms", ms_scaling_factor=0.0)
    assert dec.rank == 108
    rng = np.random.default_rng(2)
    for _ in range(20):
        e = (rng.random(hz.shape[1]) < 0.06).astype(np.uint8)
        s = ((hz @ e) % 2).astype(float)                            # the reference hands in float64 0/1
        assert np.array_equal(dec.decode(s), ref.decode(s))
        assert (dec.converge, dec.iter) == (ref.converge, ref.iter)
    with pytest.raises(ValueError):
        bposd_decoder(hz, error_rate=0.03, osd_method="no_such_osd")
    corr, conv, its = dec.decode_batch(np.zeros((3, hz.shape[0])))
    assert not corr.any() and conv.all()


# ---------------------------------------------------------------------------------------------
# statistical bar of the north star: logical error rate inside the oracle's 95 % binomial interval
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("p", [0.01, 0.02, 0.035])
def test_f32_logical_error_rate_within_oracle_confidence_interval(gpu, code225, p):
    """fp32 BP is not bit-reproducible; its LER must still fall inside the 95 % interval of the
    fp64 reference-order decoder on the same shots (and the fp64 GPU path must equal the oracle)."""
    from exp_ldpc_b200.experiment import CodeCapacityExperiment, binomial_ci
    from oracle import pyoracle as po
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 20000
    opts = dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.625)
    ref = po.run_code_capacity(code225.checks.z, code225.logicals.z, q, prior, decoder="bp", seed=8, shots=shots, **opts)
    f64 = CodeCapacityExperiment(code225, q, prior, "bp", opts, precision="f64").run(shots, seed=8)
    f32 = CodeCapacityExperiment(code225, q, prior, "bp", opts, precision="f32").run(shots, seed=8)
    assert f64["failures"] == ref["failures"] and f64["non_converged"] == ref["non_converged"]
    lo, hi = binomial_ci(ref["failures"], shots)
    assert lo <= f32["failures"] / shots <= hi, (f32["failures"], ref["failures"], lo, hi)


def test_p_sweep_rows_match_oracle_counts(gpu, code225):
    """p_sweep (python/qldpc/misc/p_sweep.py:17-40) on one GPU: per-p failure counts equal the oracle's."""
    from exp_ldpc_b200.experiment import depolarizing_noise, p_sweep
    from oracle import pyoracle as po
    opts = dict(max_iter=20, bp_method="ms", ms_scaling_factor=0.0)
    ps = [0.01, 0.03]
    df = p_sweep(3000, ps, depolarizing_noise, lambda p: {"p": p, "pm": p}, lambda p, x, z: 2 * p / 3,
                 lambda p, x, z: 2 * p / 3, code=code225, rounds=0, decoder_mode="bp", bp_osd_options=opts, seed=4)
    for row, p in zip(df.itertuples(), ps):
        q = 2 * p / 3 + p - 4 / 3 * p * p
        ref = po.run_code_capacity(code225.checks.z, code225.logicals.z, q, 2 * p / 3, decoder="bp", seed=4, shots=3000, **opts)
        assert (row.failures, row.samples, row.non_converged) == (ref["failures"], 3000, ref["non_converged"])


# ---------------------------------------------------------------------------------------------
# the multi-round readout_correction recipes (_experiment.py:12-126) on phenomenological faults
# ---------------------------------------------------------------------------------------------
def _oracle_recipe(code, rounds, data_prior, meas_prior, mode, opts, faults=None, records=None):
    """Literal per-shot restatement of BPOSDCorrect / BPOSDHybridCorrect / BPOSDCorrectSingleShot
    (_experiment.py:43-60,80-83,115-126) on top of the oracle decoders (OSD = OSD-0)."""
    from exp_ldpc_b200.spacetime import SpacetimeCode, SpacetimeCodeSingleShot
    from oracle import pyoracle as po
    hz = code.checks.z
    m, n = hz.shape
    T = rounds
    st = SpacetimeCode(hz, T)
    h_st = st.spacetime_check_matrix
    prior_st = np.concatenate([np.full((T + 1) * n, data_prior), np.full(T * m, meas_prior)])
    hd = hz.toarray().astype(np.int64)
    kw = dict(max_iter=opts["max_iter"], bp_method=opts["bp_method"], ms_scaling_factor=opts["ms_scaling_factor"])
    if mode in ("bp", "bposd"):
        dec_st = (po.OracleBP if mode == "bp" else po.OracleBPOSD0)(h_st, channel_probs=prior_st, **kw)
    elif mode == "bposd_hybrid":
        dec_st = po.OracleBP(h_st, channel_probs=prior_st, **kw)
        dec_h = po.OracleBPOSD0(hz, error_rate=data_prior, **kw)
    else:
        ss = SpacetimeCodeSingleShot(hz)
        dec_ss = po.OracleBPOSD0(ss.spacetime_check_matrix, channel_probs=np.concatenate([np.full(n, data_prior), np.full(m, meas_prior)]), **kw)
        dec_h = po.OracleBPOSD0(hz, error_rate=data_prior, **kw)
    corrs, fails = [], []
    lz = np.asarray(code.logicals.z).astype(np.int64) & 1
    mx = code.checks.x.shape[0]
    for item in (faults if records is None else records):
        if records is None:
            f = item.astype(np.int64)
            # what the reference would slice out of the measurement record
            data = f[:(T + 1) * n].reshape(T + 1, n)
            meas = f[(T + 1) * n:].reshape(T, m)
            cum = np.cumsum(data, axis=0) % 2
            readout = cum[T]
            history = lambda t: (hd @ cum[t] + meas[t]) % 2                      # noqa: E731
        else:
            rec = item.astype(np.int64)                                          # storage_sim.py:187-196
            readout = rec[(mx + m) * T:(mx + m) * T + n]
            history = lambda t: rec[(mx + m) * t + mx:(mx + m) * t + mx + m]     # noqa: E731
        if mode in ("bp", "bposd", "bposd_hybrid"):
            syndrome = st.syndrome_from_history(history, readout)
            if records is None:
                assert np.array_equal(syndrome.astype(np.int64) % 2, (h_st @ f) % 2)
            first = st.final_correction(dec_st.decode(syndrome))
            if mode == "bposd_hybrid":
                r2 = (first + readout) % 2
                final = dec_h.decode((hd @ r2) % 2)
                corr = (final + first) % 2
            else:
                corr = first
        else:
            acc = np.zeros(n, dtype=np.int64)
            for t in range(T):
                syndrome = ((hd @ acc) % 2 + history(t)) % 2
                acc = (acc + ss.final_correction(dec_ss.decode(syndrome))) % 2
            r2 = (acc + readout) % 2
            final = dec_h.decode((hd @ r2) % 2)
            corr = (final + acc) % 2
        corrs.append(corr)
        fails.append(int(((lz @ ((readout + corr) % 2)) % 2).any()))
    return np.array(corrs, dtype=np.uint8), np.array(fails, dtype=np.uint8)


@pytest.mark.parametrize("mode", ["bp", "bposd", "bposd_hybrid", "bposd_single_shot"])
@pytest.mark.parametrize("rounds", [1, 3])
def test_storage_recipes_against_oracle_restatement(gpu, code225, mode, rounds):
    from exp_ldpc_b200.experiment import StorageExperiment
    from oracle import pyoracle as po
    opts = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd0")
    data_prior, meas_prior = 0.012, 0.015
    shots = 96
    exp = StorageExperiment(code225, rounds, data_prior, meas_prior, mode, opts)
    got = exp.run(shots, seed=31, per_shot=True)
    faults = po.sample_bits(31, 0, 0, shots, po.thresholds(exp.prior_st))
    assert np.array_equal(got["faults"], faults)
    corr, fail = _oracle_recipe(code225, rounds, data_prior, meas_prior, mode, opts, faults)
    assert np.array_equal(got["corrections"], corr), mode
    assert np.array_equal(got["fail"], fail)
    assert got["failures"] == int(fail.sum())


def test_run_simulation_multi_round_modes(gpu, code225):
    """run_simulation(rounds > 0, decoder_mode) mirrors _experiment.py:154-210: circuit under the noise
    model -> record -> recipe -> list[bool]."""
    from exp_ldpc_b200.experiment import (StorageExperiment, build_storage_simulation, circuit_noise, depolarizing_noise,
                                          run_simulation)
    opts = dict(max_iter=10, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd0", osd_order=0)
    for mode, nm, args in (("bposd", depolarizing_noise, {"p": 0.018, "pm": 0.018}),
                           ("bposd_hybrid", circuit_noise, {"p": 0.002, "pm": 0.004}),
                           ("bposd_single_shot", depolarizing_noise, {"p": 0.02, "pm": 0.01})):
        out = run_simulation(64, code225, lambda x, z: 0.01, lambda x, z: 0.012, nm, args, opts, 2, mode, seed=3)
        circuit = build_storage_simulation(2, nm(**args), code225, use_x_logicals=False).circuit
        ref = StorageExperiment(code225, 2, 0.012, 0.01, mode, opts, circuit=circuit).run(64, seed=3, per_shot=True)
        assert isinstance(out, list) and len(out) == 64 and all(isinstance(v, bool) for v in out)
        assert out == [bool(v) for v in ref["fail"]]
    with pytest.raises(RuntimeError):
        run_simulation(8, code225, lambda x, z: 0.01, lambda x, z: 0.01, depolarizing_noise, {"p": 0.01, "pm": 0.01}, opts, 2,
                       "no_such_mode")



# ---------------------------------------------------------------------------------------------
# circuit-level sampling (Pauli frames) of the reference-generated storage-simulation circuits
# ---------------------------------------------------------------------------------------------
def _golden_circuit(name):
    import gzip
    with gzip.open(os.path.join(GOLDEN, f"storage_sim_n225_{name}.stim.gz"), "rt") as fh:
        return fh.read().split("\n")


@pytest.mark.parametrize("name", ["T0_depol", "T1_depol", "T3_depol", "T1_circuit"])
def test_circuit_sampler_bit_exact(gpu, name):
    from exp_ldpc_b200.circuits import CircuitSampler
    from oracle import pyoracle as po
    smp = CircuitSampler(_golden_circuit(name))
    p = smp.program
    for shot0, shots in ((0, 300), (10 ** 9 + 7, 65)):                      # ragged batch, 64-bit shot index
        got = smp.sample(shots, seed=12345, shot0=shot0)
        ref = po.circuit_sample(p.ops, p.thresholds, p.num_qubits, p.num_measurements, 12345, shot0, shots)
        assert got.shape == (shots, p.num_measurements) and got.dtype == bool
        assert np.array_equal(got.astype(np.uint8), ref)


@pytest.mark.parametrize("mode", ["bposd", "bposd_hybrid", "bposd_single_shot"])
@pytest.mark.parametrize("name,rounds", [("T1_circuit", 1), ("T3_depol", 3)])
def test_storage_recipes_on_circuit_level_noise(gpu, code225, mode, name, rounds):
    """The full reference pipeline of _experiment.py:154-210 on the device: noisy circuit -> record
    -> measurement_view / data_view -> readout_correction recipe -> logical flag."""
    from exp_ldpc_b200.experiment import StorageExperiment
    from oracle import pyoracle as po
    opts = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd0")
    data_prior, meas_prior = 0.03, 0.05
    shots = 64
    circuit = _golden_circuit(name)
    exp = StorageExperiment(code225, rounds, data_prior, meas_prior, mode, opts, circuit=circuit)
    got = exp.run(shots, seed=77, per_shot=True)
    p = exp.circuit_sampler.program
    records = po.circuit_sample(p.ops, p.thresholds, p.num_qubits, p.num_measurements, 77, 0, shots)
    assert np.array_equal(got["records"], records)
    corr, fail = _oracle_recipe(code225, rounds, data_prior, meas_prior, mode, opts, records=records)
    assert np.array_equal(got["corrections"], corr), mode
    assert np.array_equal(got["fail"], fail)



# ---------------------------------------------------------------------------------------------
# detector error model + decoder_mode 'bpd_detector' (_experiment.py:128-151,174,186-204)
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dem_t2(gpu, code225):
    from exp_ldpc_b200.circuits import compile_circuit, detector_error_model
    from exp_ldpc_b200.storage import build_storage_simulation, circuit_noise
    circuit = build_storage_simulation(2, circuit_noise(0.004, 0.006), code225).circuit
    prog = compile_circuit(circuit)
    return circuit, prog, detector_error_model(prog)


def test_detector_error_model_columns_are_fault_symptoms(dem_t2):
    """Every column of the model is the symptom of at least one elementary fault propagated by the oracle,
    and every sampled elementary fault's symptom is a column (or empty)."""
    from oracle import pyoracle as po
    _, prog, dem = dem_t2
    ndet, nfaults = dem.fault_check_matrix.shape
    assert ndet == len(prog.detectors) == 108 + 216 + 108 and dem.fault_map.shape == (9, nfaults)
    assert (dem.fault_priors > 0).all() and (dem.fault_priors < 0.5).all()
    cols = {(tuple(dem.fault_check_matrix[:, j].nonzero()[0]), tuple(dem.fault_map[:, j].nonzero()[0])) for j in range(nfaults)}
    assert len(cols) == nfaults                                          # merged: no duplicate symptoms
    rng = np.random.default_rng(0)
    noisy = prog.ops[(prog.ops[:, 0] >> 8) > 0]
    pick = noisy[rng.choice(len(noisy), 200, replace=False)]
    faults = []
    for code_ti, _a, _b, site in pick:
        code = int(code_ti) & 0xff
        faults.append((site, int(rng.integers(1, 4)) if code == 7 else int(rng.integers(1, 16)) if code == 8 else 0))
    rec = po.circuit_propagate_faults(prog.ops, prog.num_qubits, prog.num_measurements, faults).astype(np.int64)
    det = (rec @ dem.detector_matrix.T.toarray()) % 2
    obs = (rec @ dem.observable_matrix.T.toarray()) % 2
    for k in range(len(faults)):
        key = (tuple(np.flatnonzero(det[k])), tuple(np.flatnonzero(obs[k])))
        assert key == ((), ()) or key in cols, faults[k]


def test_detector_error_model_predicts_detector_rates(dem_t2):
    """P(detector fires) from the model = XOR-combination of the priors of the faults that touch it; the
    circuit sampler must reproduce it (this ties the sampler, the annotations and the model together)."""
    from exp_ldpc_b200.circuits import CircuitSampler
    circuit, prog, dem = dem_t2
    shots = 40000
    rec = CircuitSampler(prog).sample(shots, seed=5).astype(np.int64)
    fired = ((rec @ dem.detector_matrix.T.toarray()) % 2).mean(axis=0)
    h = dem.fault_check_matrix.tocsr()
    pred = np.array([(1 - np.prod(1 - 2 * dem.fault_priors[h.indices[h.indptr[i]:h.indptr[i + 1]]])) / 2 for i in range(h.shape[0])])
    sigma = np.sqrt(pred * (1 - pred) / shots)
    assert np.all(np.abs(fired - pred) < 5 * sigma + 2e-4)


def test_bpd_detector_mode_against_oracle(gpu, code225):
    """Measurement-step noise (depolarizing_noise): its detector error model is small enough for the
    shared-memory BP kernel; the circuit_noise model of the same code (15 k faults, 45 k edges) falls back
    to the global-memory kernel."""
    from exp_ldpc_b200.circuits import compile_circuit
    from exp_ldpc_b200.experiment import DetectorExperiment, run_simulation
    from exp_ldpc_b200.storage import build_storage_simulation, circuit_noise, depolarizing_noise
    from oracle import pyoracle as po
    circuit = build_storage_simulation(2, depolarizing_noise(0.01, 0.012), code225).circuit
    prog = compile_circuit(circuit)
    opts = dict(max_iter=15, bp_method="ms", ms_scaling_factor=0.0)
    shots = 48
    exp = DetectorExperiment(circuit, opts)
    dem = exp.dem
    out = run_simulation(shots, code225, lambda x, z: 0.01, lambda x, z: 0.01, depolarizing_noise, {"p": 0.01, "pm": 0.012},
                         opts, 2, "bpd_detector", seed=9)
    big = DetectorExperiment(build_storage_simulation(2, circuit_noise(0.004, 0.006), code225).circuit, opts)
    assert big.decoder.config()["kernel"] == "global"       # 15 k faults / 96 k edges: beyond 8 SMs' shared memory
    got = exp.run(shots, seed=9, per_shot=True)
    rec = po.circuit_sample(prog.ops, prog.thresholds, prog.num_qubits, prog.num_measurements, 9, 0, shots).astype(np.int64)
    assert np.array_equal(got["records"], rec)
    det = (rec @ dem.detector_matrix.T.toarray()) % 2
    obs = (rec @ dem.observable_matrix.T.toarray()) % 2
    assert np.array_equal(got["detectors"], det)
    ref = po.OracleBP(exp.dem.fault_check_matrix, channel_probs=exp.dem.fault_priors, **opts)
    fm = exp.dem.fault_map.toarray().astype(np.int64)
    for s in range(shots):
        fault_set = ref.decode(det[s])                                  # _experiment.py:147-151
        assert np.array_equal(got["fault_sets"][s], fault_set), s
        logicals = (obs[s] + fm @ fault_set) % 2
        assert got["fail"][s] == int(logicals.any())
    assert out == [bool(v) for v in got["fail"]]



# ---------------------------------------------------------------------------------------------
# global-memory BP kernel: the fallback for graphs the shared-memory kernels cannot take
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("name", _bp_case_names())
def test_global_memory_bp_kernel_matches_committed_golden_vectors(gpu, codes, name):
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, kernel="global")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])
    assert np.array_equal(out["converged"], g[name + "_conv"])


@pytest.mark.parametrize("name", _bp_case_names())
def test_cluster_bp_kernel_matches_committed_golden_vectors(gpu, codes, name):
    """Every golden BP case replayed with the messages split over a thread-block cluster (forced: these graphs
    would fit one SM)."""
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, kernel="cluster")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])
    assert np.array_equal(out["converged"], g[name + "_conv"])


@pytest.mark.parametrize("name", _bp_case_names())
def test_regular_cluster_bp_kernel_matches_committed_golden_vectors(gpu, codes, name):
    """Every golden BP case replayed by the thread-owned-state kernel in cluster mode (forced).  These check matrices have
    no block structure, so the row split cuts most columns: replicas with several remote edges, columns spanning three
    CTAs, mailboxes as large as the local message arrays."""
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, kernel="regular_cluster")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])
    assert np.array_equal(out["converged"], g[name + "_conv"])


@pytest.mark.parametrize("name", ["ms_sched_225", "ps_225", "ms_sched_1600"])
@pytest.mark.parametrize("config", [dict(ctas_per_sm=2, slots=1), dict(ctas_per_sm=2, slots=4), dict(ctas_per_sm=2, slots=2),
                                    dict(ctas_per_sm=4, slots=3), dict(ctas_per_sm=8, slots=2, threads=128),
                                    dict(ctas_per_sm=4, slots=1, threads=64)])
def test_regular_cluster_kernel_independent_of_cluster_geometry(gpu, codes, name, config):
    """Cluster size (read from `ctas_per_sm` in this family; 2, 4 or 8), slots and threads change where the messages live and
    which CTA updates which column, never the bits."""
    g = np.load(os.path.join(GOLDEN, "decoder_vectors.npz"))
    out = _run_gpu_case(name, codes, config=config, kernel="regular_cluster")
    assert np.array_equal(np.packbits(out["decodings"], axis=1), g[name + "_dec"])
    assert np.array_equal(out["iters"], g[name + "_iters"])
    assert np.array_equal(out["fail"], g[name + "_fail"])


def test_regular_cluster_kernel_soft_outputs_and_priors(gpu, code225):
    """Per-column priors, product-sum and min-sum, soft outputs of the single-shot call, fp32: cluster mode equals the
    single-CTA kernel bit for bit (same arithmetic, different placement)."""
    from exp_ldpc_b200.decoders import bp_decoder
    from exp_ldpc_b200.spacetime import SpacetimeCode
    rng = np.random.default_rng(5)
    h = SpacetimeCode(code225.checks.z, 1).spacetime_check_matrix
    probs = rng.uniform(0.004, 0.05, h.shape[1])
    err = (rng.random((300, h.shape[1])) < probs).astype(np.uint8)
    synd = (err @ h.T.toarray()) % 2
    for precision in ("f64", "f32"):
        for method in ("ms", "ps"):
            a = bp_decoder(h, channel_probs=probs, max_iter=12, bp_method=method, ms_scaling_factor=0.0, kernel="regular", precision=precision)
            b = bp_decoder(h, channel_probs=probs, max_iter=12, bp_method=method, ms_scaling_factor=0.0, kernel="regular_cluster",
                           precision=precision)
            assert b.config()["kernel"] == "regular_cluster" and b.config()["ctas_per_sm"] == 2
            ca, va, ia = a.decode_batch(synd)
            cb, vb, ib = b.decode_batch(synd)
            assert np.array_equal(ca, cb) and np.array_equal(va, vb) and np.array_equal(ia, ib), (precision, method)
            assert 0 < int(va.sum()) < synd.shape[0] or method == "ps"
            for s in (0, 7):
                da, db = a.decode(synd[s]), b.decode(synd[s])
                assert np.array_equal(da, db) and a.iter == b.iter and a.converge == b.converge
                assert np.array_equal(a.log_prob_ratios, b.log_prob_ratios)


def test_bposd_over_the_cluster_mode_equals_the_single_cta_kernel(gpu, code225):
    """BP + OSD on a spacetime matrix with the BP stage in cluster mode: the soft outputs that rank the columns for OSD are
    written by the CTA that owns a column, so corrections, convergence flags and iteration counts must equal the
    single-CTA run bit for bit (osd0 and the reference's default osd_cs 7)."""
    from exp_ldpc_b200.decoders import bposd_decoder
    from exp_ldpc_b200.spacetime import SpacetimeCode
    rng = np.random.default_rng(8)
    h = SpacetimeCode(code225.checks.z, 1).spacetime_check_matrix
    probs = np.full(h.shape[1], 0.03)
    err = (rng.random((200, h.shape[1])) < probs).astype(np.uint8)
    synd = (err @ h.T.toarray()) % 2
    for method, osd, order in (("ms", "osd0", 0), ("ps", "osd_cs", 7)):
        opts = dict(channel_probs=probs, max_iter=10, bp_method=method, ms_scaling_factor=0.0, osd_method=osd, osd_order=order)
        a = bposd_decoder(h, kernel="regular", **opts)
        b = bposd_decoder(h, kernel="regular_cluster", **opts)
        assert b.config()["kernel"] == "regular_cluster"
        ca, va, ia = a.decode_batch(synd)
        cb, vb, ib = b.decode_batch(synd)
        assert 0 < int(va.sum()) < synd.shape[0]                     # both the BP and the OSD branch are exercised
        assert np.array_equal(ca, cb) and np.array_equal(va, vb) and np.array_equal(ia, ib), (method, osd)
        assert np.array_equal((cb @ h.T.toarray()) % 2, synd)        # BP + OSD always returns a correction with the right syndrome


@pytest.mark.parametrize("kernel", ["auto", "cluster", "global"])
def test_large_graphs_take_the_cluster_kernel(gpu, code1600, kernel):
    """Spacetime matrix of the n = 1600 code, T = 3: 26 k edges do not fit one SM in fp64; the rows are split over a
    cluster along the rounds and decoded by the thread-owned-state kernel (automatic choice), or, forced, by the generic
    cluster kernel or out of global memory."""
    from exp_ldpc_b200.experiment import SpacetimeExperiment
    from oracle import pyoracle as po
    opts = dict(max_iter=8, bp_method="ms", ms_scaling_factor=0.0)
    exp = SpacetimeExperiment(code1600, 3, 0.004, 0.006, opts, kernel=kernel)
    assert exp.decoder.config()["kernel"] == ("regular_cluster" if kernel == "auto" else kernel)
    if kernel == "auto":
        assert exp.decoder.config()["ctas_per_sm"] == 4              # one round per CTA
    shots = 24 if kernel != "auto" else 2000                         # auto: more shots than slots in flight (refill path)
    got = exp.run(shots, seed=2, per_shot=True)
    h_st = exp.spacetime.spacetime_check_matrix
    m, n = code1600.checks.z.shape
    lz = np.asarray(code1600.logicals.z) & 1
    l_st = np.concatenate([lz] * 4 + [np.zeros((lz.shape[0], 3 * m), dtype=lz.dtype)], axis=1)
    ref = po.run_code_capacity(h_st, l_st, exp.prior, exp.prior, decoder="bp", seed=2, shots=shots, per_shot=True, **opts)
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["iters"], ref["iters"]) and np.array_equal(got["fail"], ref["fail"])


def test_dense_rows_take_a_multi_sm_kernel(gpu):
    """Row degree 300 (> 255) and product-sum with row degree > 16: beyond the shared-memory kernels."""
    from scipy import sparse
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    rng = np.random.default_rng(4)
    a = (rng.random((40, 600)) < 0.05).astype(np.uint32)
    a[0, :300] = 1
    h = sparse.csr_matrix(a)
    probs = rng.uniform(0.01, 0.05, 600)
    err = (rng.random((20, 600)) < probs).astype(np.int64)
    synd = (err @ a.T.astype(np.int64)) % 2
    for method in ("ms", "ps"):
        dec = bp_decoder(h, channel_probs=probs, max_iter=10, bp_method=method, ms_scaling_factor=0.0)
        assert dec.config()["kernel"] in ("cluster", "global")          # beyond the single-SM kernels' degree limits
        ref = po.OracleBP(h, channel_probs=probs, max_iter=10, bp_method=method, ms_scaling_factor=0.0)
        corr, conv, its = dec.decode_batch(synd)
        for s in range(synd.shape[0]):
            assert np.array_equal(corr[s], ref.decode(synd[s])), (method, s)
            assert its[s] == ref.iter and conv[s] == ref.converge


def test_global_memory_kernel_long_columns(gpu):
    """Columns of ~18 and ~60 edges (beyond the register-resident column update of the global-memory kernel), rows of
    60..200 edges: the literal serial sweeps must equal the serial decoder."""
    from scipy import sparse
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    rng = np.random.default_rng(11)
    for rows, cols, density in ((60, 200, 0.3), (120, 90, 0.5)):        # column degrees ~18 (<= 32) and ~60 (> 32)
        a = (rng.random((rows, cols)) < density).astype(np.uint32)
        a[:, 0] = 1
        a[0, :] = 1
        h = sparse.csr_matrix(a)
        probs = rng.uniform(0.01, 0.06, cols)
        err = (rng.random((24, cols)) < probs).astype(np.int64)
        synd = (err @ a.T.astype(np.int64)) % 2
        dec = bp_decoder(h, channel_probs=probs, max_iter=12, bp_method="ms", ms_scaling_factor=0.0, kernel="global")
        ref = po.OracleBP(h, channel_probs=probs, max_iter=12, bp_method="ms", ms_scaling_factor=0.0)
        corr, conv, its = dec.decode_batch(synd)
        for s in range(synd.shape[0]):
            assert np.array_equal(corr[s], ref.decode(synd[s])), (rows, cols, s)
            assert its[s] == ref.iter and conv[s] == ref.converge
        want = ref.decode(synd[3])
        assert np.array_equal(dec.decode(synd[3]), want) and np.allclose(dec.log_prob_ratios, ref.log_prob_ratios, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("which,method,osd,order,p", [("hgp225", "ms", "osd_cs", 7, 0.04), ("hgp225", "ps", "osd_cs", 10, 0.03),
                                                      ("hgp225", "ms", "osd_e", 6, 0.05), ("hgp1600", "ms", "osd_cs", 7, 0.025)])
def test_higher_order_osd_against_live_oracle(gpu, codes, which, method, osd, order, p):
    """The reference's default post-processing is osd_cs of order 7 (_experiment.py:218-219)."""
    from exp_ldpc_b200.experiment import CodeCapacityExperiment
    from oracle import pyoracle as po
    code = codes[which]
    q, prior = 2 * p / 3 + p - 4 / 3 * p * p, 2 * p / 3
    shots = 300 if which == "hgp225" else 64
    opts = dict(max_iter=16, bp_method=method, ms_scaling_factor=0.0, osd_method=osd, osd_order=order)
    exp = CodeCapacityExperiment(code, q, prior, "bposd", opts)
    assert (exp.decoder.osd_method, exp.decoder.osd_order) == (osd, order)
    got = exp.run(shots, seed=23, per_shot=True)
    ref = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder={"osd_cs": "bposd_cs", "osd_e": "bposd_e"}[osd],
                               osd_order=order, max_iter=16, bp_method=method, ms_scaling_factor=0.0, seed=23, shots=shots,
                               per_shot=True)
    zero = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bposd0", max_iter=16, bp_method=method,
                                ms_scaling_factor=0.0, seed=23, shots=shots, per_shot=True)
    assert (ref["decodings"] != zero["decodings"]).any()               # the search does change some solutions
    assert (got["decodings"] == ref["decodings"]).all()
    assert np.array_equal(got["fail"], ref["fail"]) and got["failures"] == ref["failures"]
    assert ref["decodings"].sum() <= zero["decodings"].sum()


# ---------------------------------------------------------------------------------------------
# edge cases: empty and single-shot batches, padded strides, isolated qubits, tiny row degrees
# ---------------------------------------------------------------------------------------------
def _odd_matrix():
    """7 x 12: an isolated column (no check), a column of degree 5, rows of degree 1 and 2."""
    from scipy import sparse
    a = np.zeros((7, 12), dtype=np.uint32)
    a[0, [0, 1, 2, 3]] = 1
    a[1, [1, 4]] = 1
    a[2, [5]] = 1
    a[3, [1, 6, 7, 8, 9]] = 1
    a[4, [1, 2, 10]] = 1
    a[5, [1, 3, 9, 10]] = 1
    a[6, [0, 4, 8]] = 1
    return sparse.csr_matrix(a)          # column 11 is isolated


@pytest.mark.parametrize("kernel", ["auto", "generic", "global"])
@pytest.mark.parametrize("method", ["ms", "ps"])
def test_irregular_tiny_graph_all_kernels(gpu, kernel, method):
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    h = _odd_matrix()
    probs = np.linspace(0.02, 0.3, 12)
    probs[11] = 0.7                                   # isolated qubit with prior > 1/2: decision is 1 forever
    synd = np.array([[(k >> r) & 1 for r in range(7)] for k in range(128)], dtype=np.uint8)      # every syndrome
    dec = bp_decoder(h, channel_probs=probs, max_iter=9, bp_method=method, ms_scaling_factor=0.0, kernel=kernel)
    ref = po.OracleBP(h, channel_probs=probs, max_iter=9, bp_method=method, ms_scaling_factor=0.0)
    corr, conv, its = dec.decode_batch(synd)
    for s in range(synd.shape[0]):
        want = ref.decode(synd[s])
        assert np.array_equal(corr[s], want), (kernel, method, s)
        assert its[s] == ref.iter and conv[s] == ref.converge
    assert corr[:, 11].all()


def test_empty_and_single_shot_batches(gpu, code225):
    from exp_ldpc_b200.decoders import bp_decoder, bposd_decoder, ssf_decoder
    hz, hx = code225.checks.z, code225.checks.x
    for dec in (bp_decoder(hz, error_rate=0.02, max_iter=5, bp_method="ms"),
                bposd_decoder(hz, error_rate=0.02, max_iter=5, bp_method="ms", osd_method="osd_cs", osd_order=3),
                ssf_decoder(hz, hx)):
        corr, conv, its = dec.decode_batch(np.zeros((0, hz.shape[0]), dtype=np.uint8))
        assert corr.shape == (0, hz.shape[1]) and conv.shape == (0,) and its.shape == (0,)
        corr, conv, its = dec.decode_batch(np.zeros((1, hz.shape[0]), dtype=np.uint8))
        assert not corr.any() and conv[0] == 1
        with pytest.raises(ValueError):
            dec.decode_batch(np.zeros((2, hz.shape[0] + 1), dtype=np.uint8))


def test_padded_strides_and_dirty_output_buffers(gpu, code225):
    """Strides larger than needed: padding words of the output are written as zero, inputs beyond
    ceil(m/32) words are ignored."""
    import torch
    from exp_ldpc_b200.decoders import TannerGraph, bp_decoder, pack_bits, row_words, unpack_bits
    hz = code225.checks.z
    m, n = hz.shape
    rng = np.random.default_rng(0)
    err = (rng.random((64, n)) < 0.03).astype(np.uint8)
    synd = ((err @ hz.T.toarray()) % 2).astype(np.uint8)
    dec = bp_decoder(TannerGraph(hz), error_rate=0.02, max_iter=20, bp_method="ms", ms_scaling_factor=0.0)
    base, _, _ = dec.decode_batch(synd)
    ss, cs = row_words(m) + 3, row_words(n) + 5
    packed = np.full((64, ss), 0xFFFFFFFF, dtype=np.uint32)
    packed[:, :row_words(m)] = pack_bits(synd)
    dev = torch.device("cuda:0")
    d_syn = torch.from_numpy(packed.view(np.int32)).to(dev)
    d_cor = torch.full((64, cs), -1, dtype=torch.int32, device=dev)
    dec.decode_batch_device(d_syn, ss, 64, d_cor, cs)
    torch.cuda.synchronize()
    out = d_cor.cpu().numpy().view(np.uint32)
    assert np.array_equal(unpack_bits(out[:, :row_words(n)], n), base)
    assert not out[:, (n + 31) // 32:].any()


# ---------------------------------------------------------------------------------------------
# regressions for the round-1 review
# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kernel", ["global", "auto"])
def test_host_pipeline_many_chunks_in_flight(gpu, code225, kernel, monkeypatch):
    """decode_batch over host buffers alternates chunks between two streams; the global-memory kernel's
    message scratch must not be shared between the two launches in flight."""
    from exp_ldpc_b200.decoders import bp_decoder
    from oracle import pyoracle as po
    hz = code225.checks.z
    rng = np.random.default_rng(11)
    shots = 1500
    err = (rng.random((shots, hz.shape[1])) < 0.04).astype(np.uint8)
    synd = ((err @ hz.T.toarray()) % 2).astype(np.uint8)
    opts = dict(error_rate=0.03, max_iter=20, bp_method="ms", ms_scaling_factor=0.0)
    monkeypatch.setenv("QD_HOST_CHUNK", "40")              # 38 chunks, two always in flight, grids smaller than the GPU
    dec = bp_decoder(hz, kernel=kernel, **opts)
    corr, conv, its = dec.decode_batch(synd)
    monkeypatch.delenv("QD_HOST_CHUNK")
    ref = po.OracleBP(hz, **opts)
    for s in range(shots):
        assert np.array_equal(corr[s], ref.decode(synd[s])), s
        assert its[s] == ref.iter and conv[s] == ref.converge


def test_bposd_counters_without_fail_buffer(gpu, code225):
    """qd_bposd_decode_batch with counters and the fused epilogue but no per-shot flag buffer: failures of
    unconverged shots are counted once (BP's count is replaced by OSD's, not added to)."""
    import torch
    from exp_ldpc_b200 import _lib
    from exp_ldpc_b200.decoders import Logicals, Sampler, TannerGraph, bposd_decoder, row_words
    hz = code225.checks.z
    m, n = hz.shape
    g = TannerGraph(hz)
    dec = bposd_decoder(g, error_rate=0.04, max_iter=6, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd0", osd_order=0)
    lg = Logicals(code225.logicals.z)
    shots, wn, wm = 4096, row_words(n), row_words(m)
    dev = torch.device("cuda:0")
    err = torch.zeros((shots, wn), dtype=torch.int32, device=dev)
    synd = torch.zeros((shots, wm), dtype=torch.int32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    Sampler(0.06, n).run_device(3, 0, 0, shots, err, wn, st)
    g.syndrome_batch_device(err, wn, shots, synd, wm, st)
    out = {}
    for with_fail in (True, False):
        corr = torch.zeros((shots, wn), dtype=torch.int32, device=dev)
        fail = torch.zeros(shots, dtype=torch.uint8, device=dev) if with_fail else None
        counters = torch.zeros(4, dtype=torch.int64, device=dev)
        _lib.check(_lib.lib().qd_bposd_decode_batch(dec.handle, dec.osd_handle, _lib.ptr(synd), wm, shots, _lib.ptr(corr), wn,
                                                    None, None, _lib.ptr(err), wn, lg.handle, _lib.ptr(fail), _lib.ptr(counters), st))
        torch.cuda.synchronize()
        out[with_fail] = counters.cpu().numpy().copy()
        if with_fail:
            assert int(fail.sum().item()) == int(counters[0].item())
    assert out[True][2] > 0                                   # some shots did need OSD
    assert np.array_equal(out[True], out[False])


def test_osd_rejects_more_rows_than_key_slots(gpu):
    """OSD keeps its (B, Y) pairs in the P * 8 bytes of the sort keys (P = power of two >= n): a redundant
    check matrix with more rows than that is refused at construction instead of overrunning."""
    from scipy import sparse
    from exp_ldpc_b200.decoders import bposd_decoder
    rng = np.random.default_rng(0)
    a = (rng.random((40, 16)) < 0.3).astype(np.uint32)
    a[:, 0] = 1
    with pytest.raises((RuntimeError, ValueError, NotImplementedError)):
        bposd_decoder(sparse.csr_matrix(a), error_rate=0.05, max_iter=4, bp_method="ms", osd_method="osd0", osd_order=0)
    b = (rng.random((16, 16)) < 0.3).astype(np.uint32)      # m = P: accepted
    b[:, 0] = 1
    bposd_decoder(sparse.csr_matrix(b), error_rate=0.05, max_iter=4, bp_method="ms", osd_method="osd0", osd_order=0)


def test_run_simulation_unseeded_calls_are_independent(gpu, code225):
    """Without a seed every call draws fresh shots (the reference's Stim sampler is entropy-seeded); with a
    seed the call is reproducible."""
    from functools import partial
    from exp_ldpc_b200.experiment import depolarizing_noise, run_simulation
    p = 0.06
    prior = lambda p_, *_: 2 / 3 * p_
    opts = {"max_iter": 10, "bp_method": "ms", "ms_scaling_factor": 0.0}
    args = (2000, code225, partial(prior, p), partial(prior, p), depolarizing_noise, {"p": p, "pm": p}, opts, 0, "bp")
    a, b = run_simulation(*args), run_simulation(*args)
    assert a != b and 0 < sum(a) < 2000
    assert run_simulation(*args, seed=4) == run_simulation(*args, seed=4)


# ---------------------------------------------------------------------------------------------
# 128-shot bit-sliced syndrome kernel (syndrome_kernel.cuh): ragged tiles, padded strides, rows that
# are not 16-byte aligned (no TMA staging), irregular and long rows (run-time table length), n and m that are
# not multiples of 32 — each against a dense numpy H e mod 2 and against the two older kernels
# ---------------------------------------------------------------------------------------------
def _random_sparse(rng, m, n, row_degrees):
    from scipy import sparse
    rows, cols = [], []
    for r in range(m):
        d = int(row_degrees(r))
        for c in rng.choice(n, size=min(d, n), replace=False):
            rows.append(r); cols.append(int(c))
    return sparse.csr_matrix((np.ones(len(rows), dtype=np.uint32), (rows, cols)), shape=(m, n))


@pytest.mark.parametrize("case", ["hgp225", "tiny", "ragged_words", "long_rows", "empty_rows"])
def test_syndrome128_edge_cases(gpu, code225, case, monkeypatch):
    import torch
    from exp_ldpc_b200.decoders import TannerGraph, pack_bits, row_words, unpack_bits
    rng = np.random.default_rng(11)
    if case == "hgp225":
        h = code225.checks.z
    elif case == "tiny":
        h = _random_sparse(rng, 3, 5, lambda r: 2)
    elif case == "ragged_words":
        h = _random_sparse(rng, 77, 333, lambda r: 1 + r % 9)            # row degrees 1..9: three table entries
    elif case == "long_rows":
        h = _random_sparse(rng, 40, 700, lambda r: 150 + r)              # run-time table length (Q = 48)
    else:
        h = _random_sparse(rng, 64, 96, lambda r: 0 if r % 3 == 0 else 4)
    m, n = h.shape
    hd = np.asarray(h.todense(), dtype=np.uint8) & 1
    g = TannerGraph(h)
    for shots, extra_e, extra_s, misalign in ((1, 0, 0, 0), (127, 0, 4, 0), (128, 4, 0, 0), (129, 0, 0, 0), (300, 8, 4, 0),
                                              (300, 0, 0, 1), (1000, 4, 0, 2)):
        es, ss = row_words(n) + extra_e, row_words(m) + extra_s
        e = rng.integers(0, 2, size=(shots, n)).astype(np.uint8)
        want = (e.astype(np.int64) @ hd.T.astype(np.int64)) % 2
        packed = torch.from_numpy(pack_bits(e, es).view(np.int32))
        if misalign:
            # rows that start 4 * misalign bytes off a 16-byte boundary: the kernel must not use the bulk copy
            flat = torch.zeros(shots * es + 4, dtype=torch.int32, device=gpu)
            d_e = flat[misalign:misalign + shots * es].view(shots, es)
            d_e.copy_(packed)
        else:
            d_e = packed.to(gpu)
        for env in ({}, {"QD_SYNDROME_NO_TMA": "1"}, {"QD_SYNDROME_OLD": "1"}, {"QD_SYNDROME_OLD": "1", "QD_SYNDROME_GATHER": "1"}):
            for k in ("QD_SYNDROME_NO_TMA", "QD_SYNDROME_OLD", "QD_SYNDROME_GATHER"):
                monkeypatch.delenv(k, raising=False)
            for k, v in env.items():
                monkeypatch.setenv(k, v)
            d_s = torch.full((shots, ss), -1, dtype=torch.int32, device=gpu)
            g.syndrome_batch_device(d_e, es, shots, d_s, ss)
            torch.cuda.synchronize()
            raw = d_s.cpu().numpy().view(np.uint32)
            assert np.array_equal(unpack_bits(raw, m), want), (case, shots, env)
            assert not raw[:, (m + 31) // 32:].any(), "stride padding words must be written as zeros"
            if m % 32:
                assert not (raw[:, m // 32] >> (m % 32)).any()

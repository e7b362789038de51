"""The per-shot ``readout_correction`` wrappers (``python/qldpc/misc/_experiment.py:12-151``) on a box without a GPU:
decoders of the same duck type (the CPU oracle) are plugged in through the factories, and every recipe is compared
with the literal restatement the GPU suite checks ``StorageExperiment`` against (``test_gpu_parity._oracle_recipe``)."""
import numpy as np
import pytest

from exp_ldpc_b200 import wrappers
from exp_ldpc_b200.spacetime import DetectorSpacetimeCode, SpacetimeCode
from oracle import pyoracle as po
from test_gpu_parity import _oracle_recipe

OPTS = dict(max_iter=12, bp_method="ms", ms_scaling_factor=0.0, osd_method="osd0", osd_order=0)
_bp = lambda h, osd_method=None, osd_order=None, **kw: po.OracleBP(h, **kw)                    # noqa: E731
_bposd = lambda h, osd_method="osd0", osd_order=0, **kw: po.OracleBPOSD(h, osd_method=osd_method, osd_order=osd_order, **kw)  # noqa: E731
FACTORIES = dict(bp_factory=_bp, bposd_factory=_bposd)


def _shots(code, rounds, data_prior, meas_prior, shots, seed):
    hz = code.checks.z
    m, n = hz.shape
    prior_st = np.concatenate([np.full((rounds + 1) * n, data_prior), np.full(rounds * m, meas_prior)])
    return po.sample_bits(seed, 0, 0, shots, po.thresholds(prior_st))


def _record_of(code, rounds, f):
    """readout and history(t) the reference would slice out of the record, from a phenomenological fault vector."""
    hz = code.checks.z
    m, n = hz.shape
    hd = hz.toarray().astype(np.int64)
    f = f.astype(np.int64)
    data = f[:(rounds + 1) * n].reshape(rounds + 1, n)
    meas = f[(rounds + 1) * n:].reshape(rounds, m)
    cum = np.cumsum(data, axis=0) % 2
    return cum[rounds], (lambda t: (hd @ cum[t] + meas[t]) % 2)


@pytest.mark.parametrize("mode,cls", [("bposd", wrappers.BPOSDCorrect), ("bposd_hybrid", wrappers.BPOSDHybridCorrect),
                                      ("bposd_single_shot", wrappers.BPOSDCorrectSingleShot)])
@pytest.mark.parametrize("rounds", [1, 3])
def test_recipes_equal_the_literal_restatement(code225, mode, cls, rounds):
    data_prior, meas_prior = 0.012, 0.015
    faults = _shots(code225, rounds, data_prior, meas_prior, 48, seed=31)
    want, want_fail = _oracle_recipe(code225, rounds, data_prior, meas_prior, mode, OPTS, faults)
    w = cls(code225, rounds, OPTS, (data_prior, meas_prior), **FACTORIES)
    lz = np.asarray(code225.logicals.z).astype(np.int64) & 1
    for s in range(faults.shape[0]):
        readout, history = _record_of(code225, rounds, faults[s])
        corr = w.readout_correction(history, readout)
        assert corr.shape == readout.shape and set(np.unique(corr).tolist()) <= {0, 1}
        assert np.array_equal(corr, want[s]), (mode, s)
        assert int(((lz @ ((readout + corr) % 2)) % 2).any()) == want_fail[s]      # _experiment.py:208-209
    assert want_fail.sum() < faults.shape[0]                                       # the recipes do correct something


def test_reference_prior_layout_is_the_views_of_the_reference(code225):
    """spacetime_code.py:75,86-92: the reference's views split the columns at T m, not at (T+1) n."""
    T = 2
    m, n = code225.checks.z.shape
    seen = {}

    def spy(h, channel_probs=None, **kw):
        seen["p"] = np.array(channel_probs)
        return po.OracleBP(h, channel_probs=channel_probs, max_iter=1, bp_method="ms")
    wrappers.BPOSDCorrect(code225, T, {}, (0.25, 0.125), bp_factory=spy, bposd_factory=spy, reference_prior_layout=True)
    assert (seen["p"][:T * m] == 0.25).all() and (seen["p"][T * m:] == 0.125).all()
    wrappers.BPOSDCorrect(code225, T, {}, (0.25, 0.125), bp_factory=spy, bposd_factory=spy)
    assert (seen["p"][:(T + 1) * n] == 0.25).all() and (seen["p"][(T + 1) * n:] == 0.125).all()
    assert seen["p"].shape == (SpacetimeCode(code225.checks.z, T).spacetime_check_matrix.shape[1],)


def test_detector_wrapper_corrects_the_observables():
    dem = "error(0.01) D0 D1 L0\nerror(0.02) D1 D2\nerror(0.015) D2 L1\nerror(0.01) D0\n"
    w = wrappers.BPDetectorCorrect(dem, dict(max_iter=10, bp_method="ms", ms_scaling_factor=0.0), bp_factory=_bp)
    code = DetectorSpacetimeCode(dem)
    fc, fm = code.fault_check_matrix.toarray().astype(np.int64), code.fault_map.toarray().astype(np.int64)
    for fault in range(4):                                   # a single fault: detectors fire, BP names it, observables clear
        f = np.zeros(4, dtype=np.int64)
        f[fault] = 1
        detector_string = np.concatenate([(fc @ f) % 2, (fm @ f) % 2])
        assert not w.readout_correction(detector_string).any()
    quiet = np.array([0, 0, 0, 1, 0])                        # no detector fired: the observable flip stays
    assert np.array_equal(w.readout_correction(quiet), [1, 0])
    again = wrappers.BPDetectorCorrect(code, dict(max_iter=10, bp_method="ms", ms_scaling_factor=0.0), bp_factory=_bp)
    assert np.array_equal(again.readout_correction(quiet), [1, 0])


def test_default_factories_are_the_gpu_decoders():
    from exp_ldpc_b200 import decoders
    assert wrappers._gpu_factories(None, None) == (decoders.bp_decoder, decoders.bposd_decoder)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,cls", [("bposd", wrappers.BPOSDCorrect), ("bposd_hybrid", wrappers.BPOSDHybridCorrect),
                                      ("bposd_single_shot", wrappers.BPOSDCorrectSingleShot)])
def test_recipes_on_the_gpu_decoders(built, code225, mode, cls):
    """Default factories: ``decoders.bp_decoder`` / ``bposd_decoder`` (one shot per call through the C ABI)."""
    import torch
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    rounds, data_prior, meas_prior = 2, 0.012, 0.015
    faults = _shots(code225, rounds, data_prior, meas_prior, 12, seed=5)
    want, _ = _oracle_recipe(code225, rounds, data_prior, meas_prior, mode, OPTS, faults)
    w = cls(code225, rounds, OPTS, (data_prior, meas_prior))
    for s in range(faults.shape[0]):
        readout, history = _record_of(code225, rounds, faults[s])
        assert np.array_equal(w.readout_correction(history, readout), want[s]), (mode, s)


def test_no_cpu_fallback_behind_the_wrappers(built, code225):
    """Without a CUDA device the default (GPU) decoders refuse loudly; nothing decodes on the host instead."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    from exp_ldpc_b200._lib import QdError
    with pytest.raises(QdError, match="no usable CUDA device"):
        wrappers.BPOSDHybridCorrect(code225, 1, OPTS, (0.01, 0.01))
    with pytest.raises(QdError, match="no usable CUDA device"):
        wrappers.BPDetectorCorrect("error(0.1) D0 L0\n", OPTS)

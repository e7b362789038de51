"""Per-shot ``readout_correction`` objects with the reference's names (``python/qldpc/misc/_experiment.py:12-151``).

The batched device pipelines (:class:`experiment.StorageExperiment`, :class:`experiment.DetectorExperiment`)
are what ``run_simulation`` runs; these classes are the drop-ins for code that drives the reference's wrapper
seam itself — one shot per call, ``readout_correction(history, data_readout) -> correction`` — with the GPU
decoder objects of :mod:`exp_ldpc_b200.decoders` behind ``.decode`` (no CPU fallback: constructing one without
the CUDA library raises).  ``bp_factory`` / ``bposd_factory`` exist so that the recipes can be checked on a box
without a GPU against decoders of the same duck type (the tests plug in the CPU oracle).

Two deviations, both deliberate and the same as in ``StorageExperiment``:

* priors are laid out by the true column split of the spacetime matrix (data faults first, ``(T+1) n`` columns,
  then ``T m`` measurement faults); the reference writes them through views that split at ``T m``
  (``spacetime_code.py:75,86-92``) — pass ``reference_prior_layout=True`` to get that;
* ``BPOSDCorrect`` hands the priors to the decoder as ``channel_probs``; the reference's keyword there is
  ``channel_prior`` (``_experiment.py:77``), which upstream ignores or rejects.
"""
from __future__ import annotations

from typing import Callable, Dict, Tuple

import numpy as np

from .spacetime import DetectorSpacetimeCode, SpacetimeCode, SpacetimeCodeSingleShot

__all__ = ["BPOSDCorrect", "BPOSDHybridCorrect", "BPOSDCorrectSingleShot", "BPDetectorCorrect"]


def _gpu_factories(bp_factory, bposd_factory):
    if bp_factory is None or bposd_factory is None:
        from . import decoders
        bp_factory = bp_factory or decoders.bp_decoder
        bposd_factory = bposd_factory or decoders.bposd_decoder
    return bp_factory, bposd_factory


def _bits(x) -> np.ndarray:
    return np.asarray(x).astype(np.int64) & 1


def _spacetime_priors(st, data_prior: float, measurement_prior: float, reference_layout: bool) -> np.ndarray:
    prior = np.zeros(st.spacetime_check_matrix.shape[1])
    if reference_layout:
        st.data_bits(prior)[:] = data_prior                           # _experiment.py:74-76, 106-108
        st.measurement_bits(prior)[:] = measurement_prior
    else:
        st.true_data_bits(prior)[:] = data_prior
        st.true_measurement_bits(prior)[:] = measurement_prior
    return prior


class BPOSDCorrect:
    """BP+OSD on the spacetime matrix, then the fold (``_experiment.py:62-83``)."""

    def __init__(self, code, rounds: int, bp_osd_options: Dict, priors: Tuple[float, float], *,
                 reference_prior_layout: bool = False, bp_factory=None, bposd_factory=None):
        _, bposd = _gpu_factories(bp_factory, bposd_factory)
        data_prior, measurement_prior = priors
        self._checks = code.checks.z
        self._spacetime_code = SpacetimeCode(self._checks, rounds)
        prior = _spacetime_priors(self._spacetime_code, data_prior, measurement_prior, reference_prior_layout)
        self._bpd = bposd(self._spacetime_code.spacetime_check_matrix, channel_probs=prior, **bp_osd_options)

    def readout_correction(self, history: Callable[[int], np.ndarray], readout) -> np.ndarray:
        syndrome = self._spacetime_code.syndrome_from_history(history, readout)
        return _bits(self._spacetime_code.final_correction(self._bpd.decode(syndrome)))


class BPOSDHybridCorrect:
    """BP on the spacetime matrix, fold, correct the readout, BP+OSD on H for what is left
    (``_experiment.py:85-126``)."""

    def __init__(self, code, rounds: int, bp_osd_options: Dict, priors: Tuple[float, float], *,
                 reference_prior_layout: bool = False, bp_factory=None, bposd_factory=None):
        bp, bposd = _gpu_factories(bp_factory, bposd_factory)
        data_prior, measurement_prior = priors
        self._rounds = rounds
        self._checks = code.checks.z
        self._bpd_final_round = bposd(self._checks, error_rate=data_prior, **bp_osd_options)
        self._spacetime_code = SpacetimeCode(self._checks, rounds)
        prior = _spacetime_priors(self._spacetime_code, data_prior, measurement_prior, reference_prior_layout)
        self._bpd = bp(self._spacetime_code.spacetime_check_matrix, channel_probs=prior, **bp_osd_options)

    def readout_correction(self, history: Callable[[int], np.ndarray], data_readout) -> np.ndarray:
        syndrome = self._spacetime_code.syndrome_from_history(history, data_readout)
        first = _bits(self._spacetime_code.final_correction(self._bpd.decode(syndrome)))
        corrected = first ^ _bits(data_readout)
        final = _bits(self._bpd_final_round.decode(_bits(self._checks @ corrected)))
        return final ^ first


class BPOSDCorrectSingleShot:
    """Round by round: BP+OSD on ``[H | I]`` with the syndrome of the running correction added; then BP+OSD
    on H for the corrected readout (``_experiment.py:12-60``)."""

    def __init__(self, code, rounds: int, bp_osd_options: Dict, priors: Tuple[float, float], *,
                 bp_factory=None, bposd_factory=None):
        _, bposd = _gpu_factories(bp_factory, bposd_factory)
        data_prior, measurement_prior = priors
        self._rounds = rounds
        self._checks = code.checks.z
        self._bpd_final_round = bposd(self._checks, error_rate=data_prior, **bp_osd_options)
        self._spacetime_code = SpacetimeCodeSingleShot(self._checks)
        prior = np.zeros(self._spacetime_code.spacetime_check_matrix.shape[1])
        self._spacetime_code.data_bits(prior)[:] = data_prior
        self._spacetime_code.measurement_bits(prior)[:] = measurement_prior
        self._bpd_single_shot = bposd(self._spacetime_code.spacetime_check_matrix, channel_probs=prior, **bp_osd_options)

    def readout_correction(self, history: Callable[[int], np.ndarray], data_readout) -> np.ndarray:
        data_readout = _bits(data_readout)
        accumulated = np.zeros_like(data_readout)
        for t in range(self._rounds):
            syndrome = _bits(self._checks @ accumulated) ^ _bits(history(t))
            accumulated ^= _bits(self._spacetime_code.final_correction(self._bpd_single_shot.decode(syndrome)))
        final = _bits(self._bpd_final_round.decode(_bits(self._checks @ (accumulated ^ data_readout))))
        return final ^ accumulated


class BPDetectorCorrect:
    """BP on the fault check matrix of a detector error model; returns the corrected logical observables
    (``_experiment.py:128-151``).  ``detector_error_model``: Stim ``.dem`` text, a ``stim.DetectorErrorModel``,
    :class:`circuits.DetectorErrorModel` or a :class:`spacetime.DetectorSpacetimeCode`."""

    def __init__(self, detector_error_model, bp_osd_options: Dict, *, bp_factory=None):
        bp, _ = _gpu_factories(bp_factory, bp_factory)
        code = detector_error_model if isinstance(detector_error_model, DetectorSpacetimeCode) else \
            DetectorSpacetimeCode(detector_error_model)
        self._detector_spacetime_code = code
        self._bpd = bp(code.fault_check_matrix, channel_probs=code.fault_priors, **bp_osd_options)

    def readout_correction(self, detector_string) -> np.ndarray:
        num_detectors = self._detector_spacetime_code.fault_check_matrix.shape[0]
        detector_string = _bits(detector_string)
        fault_set = _bits(self._bpd.decode(detector_string[:num_detectors]))
        return detector_string[num_detectors:] ^ _bits(self._detector_spacetime_code.fault_map @ fault_set)

"""The batch seam: ``run_simulation`` and ``p_sweep`` on the GPU.

Host-side mirror of ``python/qldpc/misc/_experiment.py:154-210`` (``run_simulation``) and
``python/qldpc/misc/p_sweep.py:17-40`` (``p_sweep``): same positional/keyword arguments, same
return types (``list[bool]`` of per-shot logical failures; a pandas DataFrame with the reference's
columns), with

* the Stim sampler replaced by the Philox code-capacity sampler (``rounds = 0``) or the Pauli-frame
  circuit sampler (``rounds > 0``) on the device;
* the per-shot Python loop replaced by ``qd_code_capacity_run``
  (sample -> syndrome -> decode -> residual -> logical flag, all on the GPU);
* ``multiprocessing.Pool(cpu_count())`` + ``sum`` replaced by one rank per GPU and a single
  ``all_reduce(SUM)`` of ``int64[4] = {failures, shots, non_converged, iterations}`` per p-value.

Supported ``decoder_mode``: ``'bp'`` (the first, pure-BP stage of the reference's
``bposd_hybrid``), ``'bposd'``, ``'bposd_hybrid'``, ``'bposd_single_shot'`` (``osd_method`` ``osd0``,
``osd_e`` or ``osd_cs`` of any order, including the reference's default ``osd_cs`` of order 7) and
``'ssf'`` (rounds = 0).

Randomness.  Shots are drawn from a counter-based generator (Philox4x32-10 keyed by ``seed``, counter =
global shot index ``shot0 + i``), so a ``(seed, shot0)`` pair names a fixed set of shots whatever the
number of ranks.  ``p_sweep`` passes each rank its own ``shot0``.  ``run_simulation`` called WITHOUT
``seed`` / ``shot0`` draws a fresh seed per call from the OS, like the reference (whose Stim sampler is
seeded from entropy, ``_experiment.py:172``): repeated calls give independent shots and may be summed.
Pass ``seed`` (and ``shot0``) to reproduce a run.
For ``rounds > 0`` the noisy storage circuit is built (``exp_ldpc_b200.storage``) and sampled with Pauli
frames on the device (:class:`StorageExperiment`); ``'bpd_detector'`` decodes on the circuit's detector
error model (:class:`DetectorExperiment`).
"""
from __future__ import annotations

import ctypes as C
import math
import time
from typing import Callable, Dict, Optional

import numpy as np

from . import _lib
from .codes import QuantumCode
from .decoders import Logicals, Sampler, TannerGraph, bp_decoder, bposd_decoder, row_words, ssf_decoder, unpack_bits

__all__ = ["depolarizing_noise", "circuit_noise", "trivial_noise", "NoiseModel", "build_storage_simulation", "CodeCapacityExperiment", "SpacetimeExperiment", "StorageExperiment", "DetectorExperiment", "run_simulation", "p_sweep",
           "shard_range", "binomial_ci"]


# the noise models are the reference's text rewriters (exp_ldpc_b200.storage); for rounds = 0 their closed-form
# code-capacity flip probability replaces the circuit
from .storage import (NoiseRewriter as NoiseModel, build_storage_simulation, circuit_noise,  # noqa: E402,F401
                      depolarizing_noise, trivial_noise)


def shard_range(total: int, rank: int, world: int):
    """Rank r decodes global shots [r*ceil(N/R), min(N, (r+1)*ceil(N/R)))."""
    per = math.ceil(total / world) if world > 0 else total
    lo = min(total, rank * per)
    hi = min(total, lo + per)
    return lo, hi


def binomial_ci(failures: int, shots: int, z: float = 1.959963984540054):
    """Wilson 95 % interval for a failure probability."""
    if shots == 0:
        return 0.0, 1.0
    ph = failures / shots
    den = 1.0 + z * z / shots
    mid = (ph + z * z / (2 * shots)) / den
    half = z * math.sqrt(ph * (1 - ph) / shots + z * z / (4.0 * shots * shots)) / den
    return max(0.0, mid - half), min(1.0, mid + half)


class CodeCapacityExperiment:
    """Device pipeline for one (check matrix, noise, decoder) point: owns the graph, sampler, logicals,
    decoder.  ``flip_probability`` and ``prior`` may be scalars or per-column vectors."""

    def __init__(self, code: QuantumCode, flip_probability, prior, decoder_mode: str = "bp",
                 bp_options: Optional[Dict] = None, device: int = 0, precision: str = "f64", kernel: str = "auto",
                 check_matrix=None, logicals=None):
        self.code = code
        self.device = int(device)
        hz = code.checks.z if check_matrix is None else check_matrix
        lz = code.logicals.z if logicals is None else logicals
        self.n = int(hz.shape[1])
        self.graph = TannerGraph(hz, device)
        self.logicals = Logicals(lz, device)
        self.sampler = Sampler(flip_probability, self.n, device)
        self.decoder_mode = decoder_mode
        opts = dict(bp_options or {})
        self._osd = None
        if decoder_mode in ("bp", "bposd"):
            # "bposd": the reference's default mode for rounds = 0 — BP + OSD on the code's check matrix
            # (BPOSDCorrect, _experiment.py:62-83); only osd_method "osd0" exists here
            pr = np.broadcast_to(np.asarray(prior, dtype=np.float64), (self.n,))
            kw = dict(channel_probs=pr, max_iter=opts.get("max_iter", 0), bp_method=opts.get("bp_method", "ps"),
                      ms_scaling_factor=opts.get("ms_scaling_factor", 0.0), precision=precision, kernel=kernel)
            if decoder_mode == "bposd":
                self.decoder = bposd_decoder(self.graph, osd_method=opts.get("osd_method", "osd0"),
                                             osd_order=opts.get("osd_order", 0), **kw)
                self._osd = self.decoder.osd_handle
            else:
                self.decoder = bp_decoder(self.graph, **kw)
            self._bp, self._ssf = self.decoder.handle, None
        elif decoder_mode == "ssf":
            if check_matrix is not None:
                raise NotImplementedError("small-set-flip decodes the code's own check matrix only")
            self.decoder = ssf_decoder(self.graph, code.checks.x, max_rounds=opts.get("max_rounds", 0))
            self._bp, self._ssf = None, self.decoder.handle
        elif decoder_mode in ("bposd_single_shot", "bposd_hybrid", "bpd_detector"):
            raise NotImplementedError(
                f"decoder_mode '{decoder_mode}' needs the multi-round circuit sampler / Stim DEM (SURVEY.md §8 f2)")
        else:
            raise RuntimeError("Unknown decoder operation mode")   # _experiment.py:189

    def run(self, shots: int, seed: int = 0, shot0: int = 0, per_shot: bool = False) -> dict:
        counters = np.zeros(4, dtype=np.int64)
        corr = fail = conv = its = None
        cs = row_words(self.n)
        if per_shot:
            corr = np.zeros((shots, cs), dtype=np.uint32)
            fail = np.zeros(shots, dtype=np.uint8)
            conv = np.zeros(shots, dtype=np.uint8)
            its = np.zeros(shots, dtype=np.int32)
        _lib.check(_lib.lib().qd_code_capacity_run(
            self.graph.handle, self._bp, self._ssf, self._osd, self.sampler.handle, self.logicals.handle, int(seed), int(shot0),
            int(shots), _lib.ptr(counters), _lib.ptr(corr), cs, _lib.ptr(fail), _lib.ptr(conv), _lib.ptr(its)))
        out = dict(failures=int(counters[0]), shots=int(counters[1]), non_converged=int(counters[2]),
                   iters_total=int(counters[3]))
        if per_shot:
            out.update(decodings=unpack_bits(corr, self.n), fail=fail, converged=conv, iters=its)
        return out


class SpacetimeExperiment(CodeCapacityExperiment):
    """Phenomenological multi-round decoding on the device (BASELINE.json configs[4] with the fault
    model of the decoder's own priors; Stim circuit-level sampling is SURVEY.md §8 f2).

    The decoding problem is the reference's ``SpacetimeCode(Hz, rounds)`` matrix
    (``spacetime_code.py:46-75``): columns = data faults of rounds 0..T, then T*m measurement
    faults.  A shot draws every column j with probability ``prior_j`` (``data_prior`` on data
    columns, ``meas_prior`` on measurement columns — explicit per column, not through the
    reference's mis-sized ``data_bits`` view), forms ``s = H_st f``, decodes, and
    fails iff ``Lz . fold(f ^ f_hat) != 0`` with ``fold`` = ``final_correction``
    (``spacetime_code.py:81-84``).  The fold is linear, so it is absorbed into the logical operators:
    ``L_st = [Lz | Lz | ... | Lz | 0]`` and the fused epilogue of the decode kernel does the rest."""

    def __init__(self, code: QuantumCode, rounds: int, data_prior: float, meas_prior: float,
                 bp_options: Optional[Dict] = None, device: int = 0, precision: str = "f64", kernel: str = "auto"):
        from .spacetime import SpacetimeCode
        st = SpacetimeCode(code.checks.z, rounds)
        h_st = st.spacetime_check_matrix
        m, n = code.checks.z.shape
        T = int(rounds)
        prior = np.concatenate([np.full((T + 1) * n, float(data_prior)), np.full(T * m, float(meas_prior))])
        lz = np.asarray(code.logicals.z) & 1
        l_st = np.concatenate([lz] * (T + 1) + [np.zeros((lz.shape[0], T * m), dtype=lz.dtype)], axis=1)
        self.spacetime = st
        self.rounds = T
        self.prior = prior
        super().__init__(code, prior, prior, "bp", bp_options, device=device, precision=precision, kernel=kernel,
                         check_matrix=h_st, logicals=l_st)

    def final_corrections(self, spacetime_decodings: np.ndarray) -> np.ndarray:
        """``SpacetimeCode.final_correction`` for a batch of decoded fault vectors (rows)."""
        n = self.code.num_qubits
        x = np.asarray(spacetime_decodings)[:, :(self.rounds + 1) * n].reshape(-1, self.rounds + 1, n)
        return (x.sum(axis=1) % 2).astype(np.uint8)


def syndrome_extraction_steps(code):
    """``(x_steps, z_steps)`` the reference hands to the prior callables: the larger of the maximum row
    and column weight of the X / Z check matrix (``_experiment.py:160-161``)."""
    checks = code.checks
    x_steps = int(max(checks.x.sum(axis=0).max(), checks.x.sum(axis=1).max()))
    z_steps = int(max(checks.z.sum(axis=0).max(), checks.z.sum(axis=1).max()))
    return x_steps, z_steps


def fresh_seed() -> int:
    """A 63-bit seed from the OS: what an unseeded ``run_simulation`` call uses (module docstring)."""
    import os
    return int.from_bytes(os.urandom(8), "little") >> 1


def run_simulation(samples, code, meas_prior, data_prior, noise_model, noise_model_args, bp_osd_options, rounds,
                   decoder_mode, seed: Optional[int] = None, shot0: int = 0, device: int = 0, precision: str = "f64"):
    """``run_simulation`` of ``_experiment.py:154-210``: builds the storage-simulation circuit under
    ``noise_model`` (``rounds = 0``: its closed-form code-capacity equivalent), samples it on the device,
    decodes with the ``decoder_mode`` recipe.  Returns ``list[bool]``: the logical-failure flag of every shot.

    ``seed=None`` (the default) draws a fresh seed per call, so repeated calls are independent as in the
    reference; ``seed=<int>`` makes the call reproducible (shots ``shot0 .. shot0 + samples - 1`` of that stream)."""
    if seed is None:
        seed = fresh_seed()
    x_steps, z_steps = syndrome_extraction_steps(code)                          # _experiment.py:160-161
    nm = noise_model(**noise_model_args)
    mprior = meas_prior(x_steps, z_steps)
    dprior = data_prior(x_steps, z_steps)
    if decoder_mode == "bpd_detector":
        sim = build_storage_simulation(rounds, nm, code, use_x_logicals=False)
        res = DetectorExperiment(sim.circuit, bp_osd_options, device=device, precision=precision).run(
            int(samples), seed=seed, shot0=shot0, per_shot=True)
        return [bool(v) for v in res["fail"]]
    if rounds != 0:
        # _experiment.py:165-197: build the noisy storage circuit, sample it (Pauli frames on the device)
        sim = build_storage_simulation(rounds, nm, code, use_x_logicals=False)
        exp = StorageExperiment(code, rounds, dprior, mprior, decoder_mode, bp_osd_options, device=device, precision=precision,
                                circuit=sim.circuit)
        res = exp.run(int(samples), seed=seed, shot0=shot0, per_shot=True)
        return [bool(v) for v in res["fail"]]
    exp = CodeCapacityExperiment(code, nm.code_capacity_flip_probability(), dprior, decoder_mode, bp_osd_options,
                                 device=device, precision=precision)
    res = exp.run(int(samples), seed=seed, shot0=shot0, per_shot=True)
    return [bool(v) for v in res["fail"]]


def p_sweep(samples, p_values, noise_model, noise_model_args, meas_prior, data_prior, *, code, rounds, decoder_mode,
            bp_osd_options, seed: int = 0, device: Optional[int] = None, precision: str = "f64", group=None,
            shard_runner: Optional[Callable] = None):
    """``p_sweep`` of ``misc/p_sweep.py:17-40``: one row per p with the reference's columns
    (``p_ph, failures, samples, walltime, rounds, decoder_mode`` + the BP options), plus
    ``non_converged`` / ``iters_total`` / ``n_gpus``.  Under ``torch.distributed`` every rank
    decodes its own shot range and the four counters are combined by one all-reduce per p.

    ``shard_runner(p_ph, flip_probability, prior, lo, hi) -> dict(failures, shots, non_converged,
    iters_total)`` replaces the GPU pipeline for one shard; the multi-rank CPU tests inject the
    oracle there (the product default is the CUDA pipeline and nothing else)."""
    import pandas as pd
    import torch

    dist = torch.distributed
    world, rank = (dist.get_world_size(group), dist.get_rank(group)) if dist.is_available() and dist.is_initialized() else (1, 0)
    if device is None:
        device = torch.cuda.current_device() if torch.cuda.is_available() else 0
    lo, hi = shard_range(int(samples), rank, world)
    # the reference binds p into the prior callables (misc/p_sweep.py:28) and run_simulation calls them with
    # the syndrome-extraction depths (_experiment.py:160-168)
    x_steps, z_steps = syndrome_extraction_steps(code)
    rows = []
    for p_ph in p_values:
        t0 = time.time()
        nm = noise_model(**noise_model_args(p_ph))
        dprior, mprior = data_prior(p_ph, x_steps, z_steps), meas_prior(p_ph, x_steps, z_steps)
        if shard_runner is None and decoder_mode == "bpd_detector":
            sim = build_storage_simulation(rounds, nm, code, use_x_logicals=False)
            exp = DetectorExperiment(sim.circuit, bp_osd_options, device=device, precision=precision)
            res = exp.run(hi - lo, seed=seed, shot0=lo) if hi > lo else dict(failures=0, shots=0, non_converged=0, iters_total=0)
        elif shard_runner is None and rounds != 0:
            sim = build_storage_simulation(rounds, nm, code, use_x_logicals=False)
            exp = StorageExperiment(code, rounds, dprior, mprior, decoder_mode,
                                    bp_osd_options, device=device, precision=precision, circuit=sim.circuit)
            res = exp.run(hi - lo, seed=seed, shot0=lo) if hi > lo else dict(failures=0, shots=0, non_converged=0, iters_total=0)
        elif shard_runner is None:
            exp = CodeCapacityExperiment(code, nm.code_capacity_flip_probability(), dprior, decoder_mode,
                                         bp_osd_options, device=device, precision=precision)
            res = exp.run(hi - lo, seed=seed, shot0=lo)
        else:
            res = shard_runner(p_ph, nm.code_capacity_flip_probability(), dprior, lo, hi)
        on_gpu = torch.cuda.is_available() and (not dist.is_initialized() or dist.get_backend(group) == "nccl")
        counters = torch.tensor([res["failures"], res["shots"], res["non_converged"], res["iters_total"]],
                                dtype=torch.int64, device=f"cuda:{device}" if on_gpu else "cpu")
        if world > 1:
            dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
        f, s, nc, it = (int(v) for v in counters.tolist())
        row = {"p_ph": float(p_ph), "failures": f, "samples": s, "walltime": time.time() - t0, "rounds": rounds,
               "decoder_mode": decoder_mode, **bp_osd_options, "non_converged": nc, "iters_total": it, "n_gpus": world}
        rows.append(row)
    return pd.DataFrame.from_records(rows)


# ---------------------------------------------------------------------------------------------
# command line: mirror of p_sweep_main (python/qldpc/misc/p_sweep.py:43-78) and of the option helpers
# add_bposd_args / unpack_bposd_args / load_code (python/qldpc/misc/_experiment.py:213-235)
# ---------------------------------------------------------------------------------------------
_SWEEP_SPEC = None


def parse_sweep_spec(x: str):
    """``(a, b, c)`` with floats a <= b and an int c > 0 (``p_sweep.py:43-55``)."""
    import re
    global _SWEEP_SPEC
    if _SWEEP_SPEC is None:
        _SWEEP_SPEC = re.compile(r'^\s*[(](.+),(.+),(.+)[)]\s*$')
    result = _SWEEP_SPEC.match(x)
    if result is None:
        raise RuntimeError('Unable to parse sweep specification, expecting (a, b, c) where a,b : float, c : int, '
                           'a<=b, and c > 0. Ex: (0.3, 1e3, 10)')
    lower, upper, points = float(result.group(1)), float(result.group(2)), int(result.group(3))
    if points <= 0 or lower > upper:
        raise RuntimeError('Number of points non-positive or lower bound exceeded upper bound')
    return lower, upper, points


def add_bposd_args(parser):
    """Same flags as ``_experiment.py:213-219``."""
    parser.add_argument('--bposd_max_iter', type=lambda x: int(x) if x is not None else None, default=None,
                        help='Maximum number of iterations for BP. Default is the number of qubits in the code')
    parser.add_argument('--bposd_bp_method', choices=['ps', 'ms', 'msl'], default='ps',
                        help='BP method (product-sum, min-sum, min-sum log)')
    parser.add_argument('--bposd_ms_scaling_factor', type=float, default=0,
                        help='Min sum scaling factor. Use variable scaling factor method if 0')
    parser.add_argument('--bposd_osd_method', choices=['osd_e', 'osd_cs', 'osd0'], default='osd_cs', help='OSD method')
    parser.add_argument('--bposd_osd_order', type=int, default=7, help='OSD search depth')


def unpack_bposd_args(parsed_args, code) -> dict:
    """``_experiment.py:221-229``."""
    return {
        'max_iter': parsed_args.bposd_max_iter if parsed_args.bposd_max_iter is not None else code.num_qubits,
        'bp_method': parsed_args.bposd_bp_method,
        'ms_scaling_factor': parsed_args.bposd_ms_scaling_factor,
        'osd_method': parsed_args.bposd_osd_method,
        'osd_order': parsed_args.bposd_osd_order,
    }


def p_sweep_parser():
    """The reference's command line (``misc/p_sweep.py:57-65``: same flags, same defaults — ``--rounds 1``,
    ``--decoder_mode bposd``) plus what this engine adds: the modes ``bp`` and ``ssf``, ``--precision``, ``--seed``."""
    from argparse import ArgumentParser
    from pathlib import Path
    parser = ArgumentParser(description='Sweep the physical error rate for the given quantum code on the GPU(s)')
    parser.add_argument('code', type=Path)
    parser.add_argument('--samples', type=int, help='Number of samples to take', required=True)
    parser.add_argument('--p_sweep', type=parse_sweep_spec, required=True,
                        help='lower and upper bounds of the sweep + number of points in the form (lower, upper, points)')
    parser.add_argument('--rounds', type=int, default=1,
                        help='Number of rounds of syndrome extraction (0: code capacity)')
    parser.add_argument('--decoder_mode', default='bposd',
                        choices=['bposd', 'bposd_single_shot', 'bposd_hybrid', 'bpd_detector', 'bp', 'ssf'],
                        help='BP+OSD, BP+OSD (single shot), hybrid BP + (BP+OSD), BP on the detector error model; '
                             'bp: the pure-BP first stage of the hybrid; ssf: small-set-flip (rounds = 0 only)')
    parser.add_argument('--linspace', type=bool, default=False,
                        help='linearly spaced points (default: uniform spacing in log space)')
    parser.add_argument('--precision', default='f64', choices=['f64', 'f32'])
    parser.add_argument('--seed', type=int, default=0)
    add_bposd_args(parser)
    return parser


def p_sweep_main(noise_model_args, noise_model, meas_prior, data_prior, argv=None):
    """Same command line as the reference's ``p_sweep_main`` (:func:`p_sweep_parser`); prints the result CSV on
    rank 0.  Under ``torchrun`` the shots of every p are sharded over the ranks."""
    import sys

    import torch

    from .codes import load_code
    parser = p_sweep_parser()
    args = parser.parse_args(sys.argv[1:] if argv is None else argv)
    code = load_code(args.code)
    bp_osd_options = unpack_bposd_args(args, code)
    sweep = np.linspace(*args.p_sweep) if args.linspace else np.geomspace(*args.p_sweep)
    dist = torch.distributed
    import os
    if int(os.environ.get("WORLD_SIZE", "1")) > 1 and not dist.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        dist.init_process_group("nccl")
    result = p_sweep(samples=args.samples, code=code, rounds=args.rounds, noise_model=noise_model,
                     noise_model_args=noise_model_args, meas_prior=meas_prior, data_prior=data_prior, p_values=sweep,
                     decoder_mode=args.decoder_mode, bp_osd_options=bp_osd_options, seed=args.seed, precision=args.precision)
    if not dist.is_initialized() or dist.get_rank() == 0:
        result.to_csv(sys.stdout)
    if dist.is_initialized():
        dist.destroy_process_group()
    return result


# ---------------------------------------------------------------------------------------------
# multi-round decoder recipes of the reference, on the device, phenomenological faults
# ---------------------------------------------------------------------------------------------
class StorageExperiment:
    """The ``readout_correction`` recipes of ``python/qldpc/misc/_experiment.py`` as device pipelines:

    ``bp``                 BP on the spacetime matrix, fold                     (first stage of the hybrid)
    ``bposd``              BP+OSD on the spacetime matrix, fold                 (``BPOSDCorrect``, ``:62-83``)
    ``bposd_hybrid``       BP on the spacetime matrix, fold, correct the readout, BP+OSD on H
                                                                                (``BPOSDHybridCorrect``, ``:85-126``)
    ``bposd_single_shot``  for every round: BP+OSD on ``[H | I]`` with the syndrome of the running
                           correction added, then BP+OSD on H for the readout
                                                                                (``BPOSDCorrectSingleShot``, ``:12-60``)

    Faults are phenomenological: every column of ``SpacetimeCode(Hz, rounds)`` (data faults per round,
    measurement faults per round) flips independently with its prior.  From the fault vector f the
    quantities the reference slices out of a Stim record follow: data readout = XOR of the data
    blocks of f; measured syndrome of round t = XOR of the first t+1 blocks of ``H_st f`` (the
    reference differences the history, ``spacetime_code.py:113-117``; this undoes it).  Circuit-level
    sampling needs the Pauli-frame simulator.  OSD means OSD-0.

    Everything between the sampler and the failure flags runs on the GPU through the C ABI; torch
    only owns the buffers."""

    MODES = ("bp", "bposd", "bposd_hybrid", "bposd_single_shot")

    def __init__(self, code: QuantumCode, rounds: int, data_prior: float, meas_prior: float, decoder_mode: str,
                 bp_osd_options: Optional[Dict] = None, device: int = 0, precision: str = "f64", circuit=None):
        """``circuit``: Stim text of the storage simulation (``StorageSim.circuit`` of the reference for this
        code and ``rounds``): shots are then sampled from that noisy circuit on the device (Pauli frames)
        instead of from phenomenological faults; the priors only parameterise the decoders."""
        from .spacetime import SpacetimeCode, SpacetimeCodeSingleShot
        if decoder_mode not in self.MODES:
            if decoder_mode == "bpd_detector":
                raise ValueError("decoder_mode 'bpd_detector' decodes detectors, not a readout: use DetectorExperiment")
            raise RuntimeError("Unknown decoder operation mode")   # _experiment.py:189
        opts = dict(bp_osd_options or {})
        kw = dict(max_iter=opts.get("max_iter", 0), bp_method=opts.get("bp_method", "ps"),
                  ms_scaling_factor=opts.get("ms_scaling_factor", 0.0), precision=precision, device=device)
        okw = dict(osd_method=opts.get("osd_method", "osd0"), osd_order=opts.get("osd_order", 0))
        self.code, self.mode, self.device, self.T = code, decoder_mode, int(device), int(rounds)
        hz = code.checks.z
        self.m, self.n = hz.shape
        T, m, n = self.T, self.m, self.n
        self.spacetime = SpacetimeCode(hz, T)
        h_st = self.spacetime.spacetime_check_matrix
        self.prior_st = np.concatenate([np.full((T + 1) * n, float(data_prior)), np.full(T * m, float(meas_prior))])
        self.g_st = TannerGraph(h_st, device)
        self.g_h = TannerGraph(hz, device)
        self.sampler = Sampler(self.prior_st, h_st.shape[1], device)
        self.circuit_sampler = None
        if circuit is not None:
            from .circuits import CircuitSampler
            self.circuit_sampler = CircuitSampler(circuit, device)
            self.mx = code.checks.x.shape[0]
            want = T * (self.mx + m) + n                       # storage_sim.py:187-196
            if self.circuit_sampler.num_measurements != want:
                raise ValueError(f"circuit has {self.circuit_sampler.num_measurements} measurements, the storage "
                                 f"simulation of this code with {T} rounds has {want}")
        self.logicals = Logicals(code.logicals.z, device)
        self.dec_st = self.dec_h = self.dec_ss = None
        if decoder_mode == "bp":
            self.dec_st = bp_decoder(self.g_st, channel_probs=self.prior_st, **kw)
        elif decoder_mode == "bposd":
            self.dec_st = bposd_decoder(self.g_st, channel_probs=self.prior_st, **kw, **okw)
        elif decoder_mode == "bposd_hybrid":
            self.dec_st = bp_decoder(self.g_st, channel_probs=self.prior_st, **kw)
            self.dec_h = bposd_decoder(self.g_h, error_rate=float(data_prior), **kw, **okw)
        else:
            h_ss = SpacetimeCodeSingleShot(hz).spacetime_check_matrix
            self.g_ss = TannerGraph(h_ss, device)
            self.prior_ss = np.concatenate([np.full(n, float(data_prior)), np.full(m, float(meas_prior))])
            self.dec_ss = bposd_decoder(self.g_ss, channel_probs=self.prior_ss, **kw, **okw)
            self.dec_h = bposd_decoder(self.g_h, error_rate=float(data_prior), **kw, **okw)

    # ---- small helpers over packed device buffers
    def _extract(self, src, bit_offset, nbits, dst, accumulate, dst_offset=0):
        _lib.check(_lib.lib().qd_bits_copy_batch(_lib.ptr(src), src.shape[1], int(bit_offset), src.shape[0], int(nbits),
                                                 _lib.ptr(dst), dst.shape[1], int(dst_offset), int(accumulate), self._stream))

    def _fold(self, faults, dst, accumulate=False):
        """``final_correction``: XOR of the T+1 data blocks (``spacetime_code.py:81-84``)."""
        for t in range(self.T + 1):
            self._extract(faults, t * self.n, self.n, dst, accumulate or t > 0)

    def run(self, shots: int, seed: int = 0, shot0: int = 0, per_shot: bool = False) -> dict:
        import torch
        dev = torch.device(f"cuda:{self.device}")
        torch.cuda.set_device(dev)
        self._stream = torch.cuda.current_stream().cuda_stream
        T, m, n = self.T, self.m, self.n
        ncol = (T + 1) * n + T * m
        wf, ws, wn, wm = row_words(ncol), row_words((T + 1) * m), row_words(n), row_words(m)
        buf = lambda w, dt=torch.int32: torch.zeros((shots, w), dtype=dt, device=dev)   # noqa: E731
        faults, sigma, fhat = buf(wf), buf(ws), buf(wf)
        readout, corr = buf(wn), buf(wn)
        conv = torch.zeros(shots, dtype=torch.uint8, device=dev)
        its = torch.zeros(shots, dtype=torch.int32, device=dev)
        fail = torch.zeros(shots, dtype=torch.uint8, device=dev)
        st = self._stream
        records = None
        if self.circuit_sampler is None:
            self.sampler.run_device(seed, 0, shot0, shots, faults, wf, st)
            self.g_st.syndrome_batch_device(faults, wf, shots, sigma, ws, st)
            self._fold(faults, readout)                               # the data readout of the storage experiment
        else:
            # measurement record -> data readout and differenced syndrome table (spacetime_code.py:98-119)
            per_round = self.mx + m
            records = buf(row_words(self.circuit_sampler.num_measurements))
            self.circuit_sampler.sample_device(seed, shot0, shots, records, records.shape[1], st)
            self._extract(records, per_round * T, n, readout, False)                 # data_view
            for t in range(T):
                z_off = per_round * t + self.mx                                       # measurement_view(t, get_x_checks=False)
                self._extract(records, z_off, m, sigma, True, dst_offset=t * m)
                self._extract(records, z_off, m, sigma, True, dst_offset=(t + 1) * m)
            last = buf(wm)
            self.g_h.syndrome_batch_device(readout, wn, shots, last, wm, st)
            self._extract(last, 0, m, sigma, True, dst_offset=T * m)
        iters_total = torch.zeros((), dtype=torch.int64, device=dev)
        if self.mode in ("bp", "bposd"):
            self.dec_st.decode_batch_device(sigma, ws, shots, fhat, wf, conv, its, stream=st)
            self._fold(fhat, corr)
            self.logicals.fail_batch_device(readout, corr, wn, shots, fail, None, st)
            iters_total += its.sum()
        elif self.mode == "bposd_hybrid":
            self.dec_st.decode_batch_device(sigma, ws, shots, fhat, wf, conv, its, stream=st)
            iters_total += its.sum()
            first = buf(wn)
            self._fold(fhat, first)                                   # final_round_bp_correction
            self._extract(first, 0, n, readout, True)                 # data_readout ^= it
            s2, c2 = buf(wm), buf(wn)
            self.g_h.syndrome_batch_device(readout, wn, shots, s2, wm, st)
            conv2 = torch.zeros_like(conv)
            self.dec_h.decode_batch_device(s2, wm, shots, c2, wn, conv2, its, None, readout, wn, self.logicals, fail, None, st)
            iters_total += its.sum()
            corr.copy_(first)
            self._extract(c2, 0, n, corr, True)                       # final_correction + final_round_bp_correction
        else:
            acc, hist, s_t = buf(wn), buf(wm), buf(wm)
            wss = row_words(n + m)
            ehat = buf(wss)
            for t in range(T):
                self._extract(sigma, t * m, m, hist, True)            # history(t) = measured syndrome of round t
                self.g_h.syndrome_batch_device(acc, wn, shots, s_t, wm, st)
                self._extract(hist, 0, m, s_t, True)
                self.dec_ss.decode_batch_device(s_t, wm, shots, ehat, wss, conv, its, stream=st)
                iters_total += its.sum()
                self._extract(ehat, 0, n, acc, True)                  # accumulated_correction += data part
            self._extract(acc, 0, n, readout, True)
            s2, c2 = buf(wm), buf(wn)
            self.g_h.syndrome_batch_device(readout, wn, shots, s2, wm, st)
            self.dec_h.decode_batch_device(s2, wm, shots, c2, wn, conv, its, None, readout, wn, self.logicals, fail, None, st)
            iters_total += its.sum()
            corr.copy_(acc)
            self._extract(c2, 0, n, corr, True)
        torch.cuda.synchronize(dev)
        out = dict(failures=int(fail.sum().item()), shots=int(shots), non_converged=int((conv == 0).sum().item()),
                   iters_total=int(iters_total.item()))
        if per_shot:
            out.update(corrections=unpack_bits(corr.cpu().numpy().view(np.uint32), n), fail=fail.cpu().numpy(),
                       converged=conv.cpu().numpy())
            if records is None:
                out["faults"] = unpack_bits(faults.cpu().numpy().view(np.uint32), ncol)
            else:
                out["records"] = unpack_bits(records.cpu().numpy().view(np.uint32), self.circuit_sampler.num_measurements)
        return out


class DetectorExperiment:
    """``decoder_mode = 'bpd_detector'`` (``BPDetectorCorrect``, ``_experiment.py:128-151``) on the device:
    sample the noisy circuit, form the detectors, BP on the fault check matrix of the circuit's detector
    error model, and flag the shot iff a corrected logical observable is non-zero
    (``logicals ^= fault_map @ fault_set``, ``:147-151,202-204``).

    The detector error model comes from :func:`exp_ldpc_b200.circuits.detector_error_model` (fault
    propagation on the GPU) with the detector / observable ids as row indices — i.e. what the reference's
    ``DetectorSpacetimeCode`` means, not what its ``lil_to_coo`` writes (``spacetime_code.py:168-171``)."""

    def __init__(self, circuit, bp_osd_options: Optional[Dict] = None, device: int = 0, precision: str = "f64"):
        from scipy import sparse

        from .circuits import CircuitSampler, compile_circuit, detector_error_model
        opts = dict(bp_osd_options or {})
        self.device = int(device)
        self.program = compile_circuit(circuit)
        self.dem = detector_error_model(self.program, device)
        self.sampler = CircuitSampler(self.program, device)
        self.nmeas = self.program.num_measurements
        self.ndet, self.nfaults = self.dem.fault_check_matrix.shape
        self.g_det = TannerGraph(self.dem.detector_matrix, device)
        self.decoder = bp_decoder(self.dem.fault_check_matrix, channel_probs=self.dem.fault_priors,
                                  max_iter=opts.get("max_iter", 0), bp_method=opts.get("bp_method", "ps"),
                                  ms_scaling_factor=opts.get("ms_scaling_factor", 0.0), precision=precision, device=device)
        both = sparse.hstack([self.dem.observable_matrix, self.dem.fault_map]).toarray()
        self.logicals = Logicals(both, device)

    def run(self, shots: int, seed: int = 0, shot0: int = 0, per_shot: bool = False) -> dict:
        import torch
        dev = torch.device(f"cuda:{self.device}")
        torch.cuda.set_device(dev)
        st = torch.cuda.current_stream().cuda_stream
        wr, wd, wf = row_words(self.nmeas), row_words(self.ndet), row_words(self.nfaults)
        wc = row_words(self.nmeas + self.nfaults)
        buf = lambda w: torch.zeros((shots, w), dtype=torch.int32, device=dev)           # noqa: E731
        rec, det, fhat, cat = buf(wr), buf(wd), buf(wf), buf(wc)
        conv = torch.zeros(shots, dtype=torch.uint8, device=dev)
        its = torch.zeros(shots, dtype=torch.int32, device=dev)
        fail = torch.zeros(shots, dtype=torch.uint8, device=dev)
        self.sampler.sample_device(seed, shot0, shots, rec, wr, st)
        self.g_det.syndrome_batch_device(rec, wr, shots, det, wd, st)
        self.decoder.decode_batch_device(det, wd, shots, fhat, wf, conv, its, stream=st)
        copy = _lib.lib().qd_bits_copy_batch
        _lib.check(copy(_lib.ptr(rec), wr, 0, shots, self.nmeas, _lib.ptr(cat), wc, 0, 0, st))
        _lib.check(copy(_lib.ptr(fhat), wf, 0, shots, self.nfaults, _lib.ptr(cat), wc, self.nmeas, 0, st))
        self.logicals.fail_batch_device(cat, None, wc, shots, fail, None, st)
        torch.cuda.synchronize(dev)
        out = dict(failures=int(fail.sum().item()), shots=int(shots), non_converged=int((conv == 0).sum().item()),
                   iters_total=int(its.sum().item()))
        if per_shot:
            out.update(fail=fail.cpu().numpy(), converged=conv.cpu().numpy(),
                       records=unpack_bits(rec.cpu().numpy().view(np.uint32), self.nmeas),
                       detectors=unpack_bits(det.cpu().numpy().view(np.uint32), self.ndet),
                       fault_sets=unpack_bits(fhat.cpu().numpy().view(np.uint32), self.nfaults))
        return out

"""World-size-2 `gloo` test of the shard + all-reduce layer on CPU.

The reference fans shots out with `multiprocessing.Pool` and sums the per-worker flags
(python/qldpc/misc/p_sweep.py:18-32).  Here every rank runs `exp_ldpc_b200.experiment.p_sweep` on its
own shard and the four counters meet in one all-reduce.  No GPU is present in this test, so the
per-shard decode is injected through `shard_runner` and done by the CPU oracle (test infrastructure);
what is under test is the host logic: shard ranges, counter-based sampling that is independent of
the rank count, the collective, and the DataFrame the reference's `p_sweep` would have produced.
"""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPTS = dict(max_iter=16, bp_method="ms", ms_scaling_factor=0.0)
SAMPLES = 301          # odd on purpose: ragged shards
P_VALUES = [0.01, 0.03]


def _oracle_runner(code):
    from oracle import pyoracle as po

    def run(p_ph, q, prior, lo, hi):
        if hi <= lo:
            return dict(failures=0, shots=0, non_converged=0, iters_total=0)
        r = po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder="bp", seed=5, shot0=lo, shots=hi - lo,
                                 threads=1, **OPTS)
        return dict(failures=r["failures"], shots=r["shots"], non_converged=r["non_converged"], iters_total=r["iters_total"])
    return run


def _sweep(code, **kw):
    from exp_ldpc_b200.experiment import depolarizing_noise, p_sweep
    return p_sweep(SAMPLES, P_VALUES, depolarizing_noise, lambda p: {"p": p, "pm": p}, lambda p, x, z: 2 * p / 3,
                   lambda p, x, z: 2 * p / 3, code=code, rounds=0, decoder_mode="bp", bp_osd_options=OPTS, seed=5,
                   shard_runner=_oracle_runner(code), **kw)


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from exp_ldpc_b200.codes import load_code
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    code = load_code(os.path.join(ROOT, "codes", "hgp_34_n225.qecc"))
    df = _sweep(code)
    df.drop(columns=["walltime"]).to_json(os.path.join(out_dir, f"rank{rank}.json"))
    dist.barrier()
    dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 3])
def test_p_sweep_shards_and_allreduce_match_single_rank(tmp_path, code225, world):
    import pandas as pd
    import torch.multiprocessing as mp
    single = _sweep(code225).drop(columns=["walltime"])
    assert list(single["samples"]) == [SAMPLES] * len(P_VALUES)
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    for r in range(world):
        df = pd.read_json(os.path.join(str(tmp_path), f"rank{r}.json"))
        for col in ("failures", "samples", "non_converged", "iters_total"):
            assert list(df[col]) == list(single[col]), (r, col)     # every rank holds the global sums
        assert list(df["n_gpus"]) == [world] * len(P_VALUES)
        np.testing.assert_allclose(df["p_ph"], P_VALUES)


def test_reference_dataframe_columns(code225):
    df = _sweep(code225)
    # python/qldpc/misc/p_sweep.py:32-39: p_ph, failures, samples, walltime + the run options
    for col in ("p_ph", "failures", "samples", "walltime", "rounds", "decoder_mode", "max_iter", "bp_method", "ms_scaling_factor"):
        assert col in df.columns


def test_p_sweep_hands_the_extraction_depths_to_the_priors(code225):
    """The reference binds p (misc/p_sweep.py:28) and run_simulation calls prior(x_steps, z_steps) with the
    code's syndrome-extraction depths (_experiment.py:160-168): a depth-dependent prior must see them."""
    from exp_ldpc_b200.experiment import depolarizing_noise, p_sweep, syndrome_extraction_steps
    steps = syndrome_extraction_steps(code225)
    assert steps == (7, 7)                                  # (3,4) HGP: row weight 7, column weight <= 4
    seen = []

    def depth_prior(p, x_steps, z_steps):
        seen.append((p, x_steps, z_steps))
        return 0.5 - 0.5 * (1 - 2 * p) ** z_steps           # 0 for z_steps = 0: the decoder would reject it

    priors = []

    def runner(p_ph, q, prior, lo, hi):
        priors.append(prior)
        return dict(failures=0, shots=hi - lo, non_converged=0, iters_total=0)

    p_sweep(10, [0.01], depolarizing_noise, lambda p: {"p": p, "pm": p}, depth_prior, depth_prior, code=code225, rounds=0,
            decoder_mode="bp", bp_osd_options=OPTS, shard_runner=runner)
    assert seen and all(s == (0.01, 7, 7) for s in seen)
    assert priors == [pytest.approx(0.5 - 0.5 * 0.98 ** 7)]

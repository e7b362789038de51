"""Fixed-seed decoder cases shared by the oracle regression test, the GPU parity tests and
tests/golden/make_decoder_vectors.py."""
from oracle import pyoracle as po

# name -> (code key, decoder, options, p, shots, seed)
CASES = {
    "ms_sched_225": ("hgp225", "bp", dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0), 0.05, 512, 0),
    "ms_sched_225_lowp": ("hgp225", "bp", dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0), 0.015, 512, 4),
    "ms_a0625_225": ("hgp225", "bp", dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.625), 0.02, 512, 1),
    "ms_a1_225": ("hgp225", "bp", dict(max_iter=20, bp_method="ms", ms_scaling_factor=1.0), 0.02, 512, 2),
    "ps_225": ("hgp225", "bp", dict(max_iter=32, bp_method="ps", ms_scaling_factor=0.0), 0.02, 512, 0),
    "ms_sched_225_maxn": ("hgp225", "bp", dict(max_iter=0, bp_method="ms", ms_scaling_factor=0.0), 0.03, 128, 3),
    "ssf_225": ("hgp225", "ssf", dict(), 0.01, 512, 0),
    "ms_sched_1600": ("hgp1600", "bp", dict(max_iter=32, bp_method="ms", ms_scaling_factor=0.0), 0.02, 128, 0),
    "ps_1600": ("hgp1600", "bp", dict(max_iter=16, bp_method="ps", ms_scaling_factor=0.0), 0.02, 64, 1),
    "ssf_1600": ("hgp1600", "ssf", dict(), 0.008, 64, 0),
}


def case_probabilities(p):
    q = 2.0 * p / 3.0 + p - 4.0 / 3.0 * p * p      # depolarizing_noise(p, pm=p), rounds = 0
    return q, 2.0 * p / 3.0                        # sampler flip probability, decoder prior


def run_case(name, codes, threads=0):
    key, decoder, opts, p, shots, seed = CASES[name]
    code = codes[key]
    q, prior = case_probabilities(p)
    return po.run_code_capacity(code.checks.z, code.logicals.z, q, prior, decoder=decoder, hx=code.checks.x,
                                max_iter=opts.get("max_iter", 0), bp_method=opts.get("bp_method", "ms"),
                                ms_scaling_factor=opts.get("ms_scaling_factor", 0.0), seed=seed, shots=shots,
                                threads=threads, per_shot=True)

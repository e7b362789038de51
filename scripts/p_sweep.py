#!/usr/bin/env python
"""GPU counterpart of the reference's scripts/p_sweep.py (same noise model, priors and CLI):

    python scripts/p_sweep.py codes/hgp_34_n1600.qecc --samples 10000000 --p_sweep "(0.01,0.1,10)" \
        --bposd_bp_method ms --bposd_max_iter 32 --decoder_mode bp --rounds 0
    (defaults as in the reference: --rounds 1 --decoder_mode bposd, i.e. BP + osd_cs order 7 over one extraction round)
    torchrun --nproc-per-node 8 --master-addr 127.0.0.1 scripts/p_sweep.py ...      # shots sharded over 8 GPUs
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from exp_ldpc_b200.experiment import depolarizing_noise, p_sweep_main  # noqa: E402


def constant_pheno_prior(p, _, __):          # scripts/p_sweep.py:4-5 of the reference
    return 2 / 3 * p


if __name__ == '__main__':
    noise_model_args = lambda p_ph: {'p': p_ph, 'pm': p_ph}      # noqa: E731  (reference scripts/p_sweep.py:8)
    p_sweep_main(noise_model_args, depolarizing_noise, constant_pheno_prior, constant_pheno_prior)

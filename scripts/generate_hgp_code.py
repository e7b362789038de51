#!/usr/bin/env python
"""Counterpart of the reference's scripts/generate_hgp_code.py (same arguments): a random (dv, dc)-biregular
hypergraph product code as a `qecc` file plus its noiseless syndrome-extraction circuit as Stim text.

    python scripts/generate_hgp_code.py 4 3 12 --seed 0 --save_code /tmp/hgp225.qecc --save_circuit /tmp/hgp225.stim

The graph is drawn with numpy only: a seed does not reproduce the reference's networkx-derived graph (the benchmark
codes under codes/ were written by the reference's own generator, tests/golden/make_golden.py)."""
import argparse
import os
import sys
from pathlib import Path

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from exp_ldpc_b200 import biregular_hgp, build_storage_simulation, noise_model, write_quantum_code  # noqa: E402

if __name__ == '__main__':
    parser = argparse.ArgumentParser(description='Generate a hypergraph product code and associated syndrome extraction circuit')
    parser.add_argument('dc', type=int, help='Classical code check degree in input to hypergraph product code')
    parser.add_argument('dv', type=int, help='Classical code vertex degree in input to hypergraph product code')
    parser.add_argument('nv', type=int, help='Classical code data vertex count; (nv dv)/dc must be an integer. The quantum code has '
                                             '(1+(dv/dc)^2) nv^2 qubits and (2dv/dc) nv^2 checks.')
    parser.add_argument('--girth_bound', type=int, default=None,
                        help='Remove all cycles of length <= girth_bound in the classical tanner graph')
    parser.add_argument('--girth_bound_patience', type=int, default=100000, help='Timeout for random swaps to achieve girth bound')
    parser.add_argument('--rounds', type=int, default=1, help='Number of rounds of syndrome extraction')
    parser.add_argument('--seed', type=int, default=None, help='PRNG seed')
    parser.add_argument('--save_code', type=Path, help='File path to save the code to')
    parser.add_argument('--save_circuit', type=Path, help='File path to save the syndrome extraction circuit to')
    parser.add_argument('--compute_logicals', type=bool, default=True,
                        help='Optionally compute logical operators of the code, Warning: This has O(n^3) time complexity')
    args = parser.parse_args()

    for dest, what in ((args.save_code, 'Code'), (args.save_circuit, 'Circuit')):
        if dest is not None and dest.exists():
            print(f'{what} save destination already exists')
            sys.exit(-1)

    code = biregular_hgp(args.nv, args.dv, args.dc, seed=args.seed, compute_logicals=args.compute_logicals,
                         girth_bound=args.girth_bound, girth_bound_patience=args.girth_bound_patience)
    storage_sim = build_storage_simulation(args.rounds, noise_model.trivial_noise(), code, use_x_logicals=False)

    if args.save_code is not None:
        with args.save_code.open('w') as code_file:
            write_quantum_code(code_file, code)
    else:
        write_quantum_code(sys.stdout, code)
    if args.save_circuit is not None:
        with args.save_circuit.open('w') as circuit_file:
            circuit_file.write('\n'.join(storage_sim.circuit))

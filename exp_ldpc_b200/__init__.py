"""exp_ldpc_b200 — B200-native batched qLDPC decoding engine (drop-in for the qldpc decode path).

The names the reference package exports (``python/qldpc/__init__.py:1-13``) resolve here too, lazily —
``from exp_ldpc_b200 import biregular_hgp, build_storage_simulation, noise_model, write_quantum_code``
works as ``scripts/generate_hgp_code.py`` of the reference expects — next to this package's own
decoder objects.  Importing the package loads neither torch nor the CUDA library.
"""
import importlib

_EXPORTS = {
    # containers and IO  (qecc_util.py, quantum_code_io.py)
    "QuantumCode": "codes", "QuantumCodeChecks": "codes", "QuantumCodeLogicals": "codes", "make_check_matrix": "codes",
    "read_quantum_code": "codes", "write_quantum_code": "codes", "load_code": "codes",
    # code families  (hypergraph_product_code.py, qc_lifted_product_code.py, lifted_product_code.py, random_code.py)
    "biregular_hgp": "code_builders", "random_test_hgp": "code_builders", "hypergraph_product": "code_builders",
    "random_biregular_matrix": "code_builders", "remove_short_cycles": "code_builders",
    "qc_lifted_product_code": "code_builders", "lifted_product": "code_builders",
    "lifted_product_code_pgl2": "cayley_codes", "lifted_product_code_cyclic": "cayley_codes",
    "lifted_product_code": "cayley_codes", "random_check_matrix": "cayley_codes",
    # storage simulation and noise  (storage_sim.py, noise_model.py, qecc_util.py:120-155)
    "build_storage_simulation": "storage", "build_perfect_circuit": "storage", "CircuitTargets": "storage",
    "StorageSim": "storage", "schedule_checks": "storage", "edge_color_bipartite": "storage",
    # spacetime codes  (spacetime_code.py)
    "SpacetimeCode": "spacetime", "SpacetimeCodeSingleShot": "spacetime", "DetectorSpacetimeCode": "spacetime",
    "DetectorErrorModel": "circuits", "detector_error_model": "circuits", "CircuitSampler": "circuits",
    # decoders and experiments  (ldpc's duck type; misc/_experiment.py, misc/p_sweep.py)
    "bp_decoder": "decoders", "bposd_decoder": "decoders", "ssf_decoder": "decoders",
    "run_simulation": "experiment", "p_sweep": "experiment",
    "BPOSDCorrect": "wrappers", "BPOSDHybridCorrect": "wrappers", "BPOSDCorrectSingleShot": "wrappers",
    "BPDetectorCorrect": "wrappers",
}
_MODULES = {"noise_model": "storage"}        # `noise_model.trivial_noise()` etc. live in storage.py

__all__ = sorted(list(_EXPORTS) + list(_MODULES))


def __getattr__(name):
    if name in _EXPORTS:
        return getattr(importlib.import_module(f".{_EXPORTS[name]}", __name__), name)
    if name in _MODULES:
        return importlib.import_module(f".{_MODULES[name]}", __name__)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

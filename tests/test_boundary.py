"""The drop-in boundary: the C-ABI library loads on a CPU-only box and exports every symbol
include/qldpc_b200.h declares; compute entry points fail loudly (no CPU fallback) without a GPU;
the product package never touches oracle/.  CPU only."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "qldpc_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(qd_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_all_exported(built):
    from exp_ldpc_b200 import _lib
    L = ctypes.CDLL(_lib.LIB_PATH)
    declared = _declared_symbols()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/qldpc_b200.h but not exported"
    assert sorted(_lib.SIGNATURES) == declared, "ctypes signature table out of sync with the header"
    assert L.qd_abi_version() == 2


def test_exports_are_plain_c(built):
    from exp_ldpc_b200 import _lib
    out = subprocess.run(["nm", "-D", "--defined-only", _lib.LIB_PATH], capture_output=True, text=True).stdout
    exported = {ln.split()[-1] for ln in out.splitlines() if " T " in ln}
    for name in _declared_symbols():
        assert name in exported


def test_library_is_sm100a_only(built):
    from exp_ldpc_b200 import _lib
    cuobjdump = "/usr/local/cuda/bin/cuobjdump"
    if not os.path.exists(cuobjdump):
        pytest.skip("cuobjdump not available")
    out = subprocess.run([cuobjdump, "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


def test_no_cpu_fallback_without_gpu(built, code225):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from exp_ldpc_b200 import _lib
    from exp_ldpc_b200.decoders import TannerGraph, bp_decoder
    assert _lib.lib().qd_device_count() == 0
    with pytest.raises(_lib.QdError) as ei:
        TannerGraph(code225.checks.z)
    assert ei.value.code == _lib.QD_ENODEV
    with pytest.raises(_lib.QdError):
        bp_decoder(code225.checks.z, error_rate=0.01)


def test_missing_extension_fails_loudly(monkeypatch, tmp_path):
    from exp_ldpc_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="CUDA extension not built"):
        _lib.lib()


def test_product_never_imports_the_oracle():
    """No product source may import, include, link, dlopen or call anything under oracle/
    (comments that cite the oracle as documentation are stripped first)."""
    pkg = os.path.join(ROOT, "exp_ldpc_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            path = os.path.join(dirpath, f)
            if f.endswith(".py"):
                code = re.sub(r'"""(.|\n)*?"""', "", open(path).read())
                code = re.sub(r"#.*", "", code)
                assert not re.search(r"pyoracle|libqldpc_oracle|import\s+oracle\b|from\s+oracle\b|orc_", code), path
            elif f.endswith((".cu", ".cuh", ".h", ".cpp")):
                code = re.sub(r"/\*(.|\n)*?\*/", "", open(path).read())
                code = re.sub(r"//.*", "", code)
                assert not re.search(r"oracle|orc_", code), path
    # the user-facing scripts are product too (checker scripts that call the oracle live under tests/tools)
    for f in os.listdir(os.path.join(ROOT, "scripts")):
        if f.endswith(".py"):
            code = re.sub(r'"""(.|\n)*?"""', "", open(os.path.join(ROOT, "scripts", f)).read())
            code = re.sub(r"#.*", "", code)
            assert not re.search(r"pyoracle|libqldpc_oracle|import\s+oracle\b|from\s+oracle\b|orc_", code), f
    # and the built library does not link it
    from exp_ldpc_b200 import _lib
    if os.path.exists(_lib.LIB_PATH):
        out = subprocess.run(["ldd", _lib.LIB_PATH], capture_output=True, text=True).stdout
        assert "oracle" not in out


def test_pack_unpack_roundtrip():
    from exp_ldpc_b200.decoders import bernoulli_thresholds, pack_bits, row_words, unpack_bits
    rng = np.random.default_rng(0)
    for nbits in (1, 31, 32, 33, 108, 225, 1600):
        b = rng.integers(0, 2, size=(7, nbits)).astype(np.uint8)
        p = pack_bits(b)
        assert p.dtype == np.uint32 and p.shape == (7, row_words(nbits)) and p.shape[1] % 4 == 0
        assert np.array_equal(unpack_bits(p, nbits), b)
        j = nbits - 1
        assert ((p[:, j >> 5] >> (j & 31)) & 1).tolist() == b[:, j].tolist()
    from oracle import pyoracle as po
    q = np.array([0.0, 1e-12, 0.05, 0.5, 1.0])
    assert np.array_equal(bernoulli_thresholds(q), po.thresholds(q))


def test_shard_range_covers_everything():
    from exp_ldpc_b200.experiment import binomial_ci, shard_range
    for total in (0, 1, 10, 1000, 12345):
        for world in (1, 2, 3, 8):
            spans = [shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
    lo, hi = binomial_ci(50, 1000)
    assert lo < 0.05 < hi and 0.03 < lo and hi < 0.07


def test_noise_model_flip_probability():
    from exp_ldpc_b200.experiment import depolarizing_noise
    nm = depolarizing_noise(0.03, 0.03)
    assert abs(nm.code_capacity_flip_probability() - (0.02 + 0.03 - 2 * 0.02 * 0.03)) < 1e-15
    assert abs(depolarizing_noise(0.03, 0.0).code_capacity_flip_probability() - 0.02) < 1e-15


def test_p_sweep_cli_options_match_reference():
    """Flag names / defaults of python/qldpc/misc/_experiment.py:213-229 and p_sweep.py:43-67."""
    from argparse import ArgumentParser
    from exp_ldpc_b200.experiment import add_bposd_args, parse_sweep_spec, unpack_bposd_args
    ap = ArgumentParser()
    add_bposd_args(ap)
    a = ap.parse_args([])

    class _Code:
        num_qubits = 225
    o = unpack_bposd_args(a, _Code)
    assert o == {'max_iter': 225, 'bp_method': 'ps', 'ms_scaling_factor': 0, 'osd_method': 'osd_cs', 'osd_order': 7}
    a = ap.parse_args(['--bposd_max_iter', '32', '--bposd_bp_method', 'ms', '--bposd_ms_scaling_factor', '0.625'])
    assert unpack_bposd_args(a, _Code)['max_iter'] == 32
    assert parse_sweep_spec("(0.01, 1e-1, 10)") == (0.01, 0.1, 10)
    for bad in ("0.1,0.2,3", "(0.2,0.1,3)", "(0.1,0.2,0)"):
        with pytest.raises(RuntimeError):
            parse_sweep_spec(bad)


def test_load_code_takes_a_path_or_parsed_arguments():
    """_experiment.py:231-235 hands load_code the argparse namespace."""
    import os
    from argparse import Namespace
    from pathlib import Path
    from exp_ldpc_b200.codes import load_code
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "codes", "hgp_34_n225.qecc")
    a, b, c = load_code(path), load_code(Path(path)), load_code(Namespace(code=Path(path)))
    assert a.num_qubits == b.num_qubits == c.num_qubits == 225
    assert (a.checks.z != c.checks.z).nnz == 0


def test_p_sweep_command_line_defaults_match_reference():
    """misc/p_sweep.py:57-65: a bare `p_sweep.py code --samples N --p_sweep (..)` means one round of syndrome
    extraction decoded with BP+OSD (osd_cs, order 7), as in the reference."""
    from exp_ldpc_b200.experiment import p_sweep_parser
    a = p_sweep_parser().parse_args(["code.qecc", "--samples", "10", "--p_sweep", "(0.01,0.1,3)"])
    assert a.rounds == 1 and a.decoder_mode == "bposd" and a.linspace is False
    assert a.bposd_osd_method == "osd_cs" and a.bposd_osd_order == 7 and a.bposd_bp_method == "ps"
    assert str(a.code) == "code.qecc" and a.samples == 10 and a.p_sweep == (0.01, 0.1, 3)
    for mode in ("bposd", "bposd_single_shot", "bposd_hybrid", "bpd_detector", "bp", "ssf"):
        assert p_sweep_parser().parse_args(["c", "--samples", "1", "--p_sweep", "(1,1,1)", "--decoder_mode", mode]).decoder_mode == mode
    with pytest.raises(SystemExit):
        p_sweep_parser().parse_args(["c", "--samples", "1", "--p_sweep", "(1,1,1)", "--decoder_mode", "nope"])


# ---- cluster mode of the row-regular BP kernel (family 9): the host-side plan checks itself without a GPU
def _plan_check(h, ctas, banks=16):
    import ctypes as C
    from scipy import sparse
    from exp_ldpc_b200 import _lib
    h = sparse.csr_matrix(h)
    h.sort_indices()
    rp = np.ascontiguousarray(h.indptr, dtype=np.int32)
    ci = np.ascontiguousarray(h.indices, dtype=np.int32)
    stats = np.zeros(8, dtype=np.int64)
    rc = _lib.lib().qd_bp_cluster_plan_check(h.shape[0], h.shape[1], _lib.ptr(rp), _lib.ptr(ci), ctas, banks, _lib.ptr(stats))
    return rc, dict(zip(("errors", "local_edges", "mailbox", "lines", "nrec", "rows", "exports", "_"), stats.tolist()))


@pytest.mark.parametrize("rounds,ctas", [(1, 2), (3, 2), (3, 4), (2, 3), (3, 3), (7, 8)])
def test_cluster_plan_of_spacetime_matrices_is_consistent(built, code225, rounds, ctas):
    """Rows of [blockdiag(H x (T+1)) | M] (spacetime_code.py:46-75) split over the CTAs of a cluster: every edge is local to
    one CTA, every remote edge of a replica is fed by one export, all accesses are bank-conflict free; a split along the
    rounds cuts only measurement columns."""
    from exp_ldpc_b200.spacetime import SpacetimeCode
    hz = code225.checks.z
    m, n = hz.shape
    h = SpacetimeCode(hz, rounds).spacetime_check_matrix
    for banks in (16, 32):
        rc, st = _plan_check(h, ctas, banks)
        assert rc == 0 and st["errors"] == 0, st
        assert st["local_edges"] == h.nnz
        if (rounds + 1) % ctas == 0:
            assert st["mailbox"] == 2 * (ctas - 1) * m       # one measurement column per check and boundary, a replica on either side
            assert st["rows"] == (rounds + 1) // ctas * m


def test_cluster_plan_of_an_unstructured_split_is_consistent(built, code225):
    """A plain (3,4) HGP check matrix has no block structure: the split cuts most columns (replicas with several remote
    edges, columns spanning three CTAs).  Slow on the GPU, but the plan must still be right."""
    hz = code225.checks.z
    for ctas in (2, 3):
        rc, st = _plan_check(hz, ctas)
        assert rc == 0 and st["errors"] == 0, st
        assert st["local_edges"] == hz.nnz and st["mailbox"] > hz.nnz // 4


def test_cluster_plan_rejects_what_the_kernel_cannot_hold(built):
    from scipy import sparse
    from exp_ldpc_b200 import _lib
    dense_row = sparse.csr_matrix(np.ones((4, 12), dtype=np.uint8))      # rows of 12 edges > 9
    rc, _ = _plan_check(dense_row, 2)
    assert rc == _lib.QD_EUNSUP

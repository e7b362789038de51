"""DetectorSpacetimeCode from a detector error model in Stim's text format (host only; SURVEY 8(a) row 9c,
``python/qldpc/spacetime_code.py:122-183``): ids made absolute, loops unrolled, row = detector / observable id."""
import numpy as np
import pytest

from exp_ldpc_b200 import DetectorSpacetimeCode
from exp_ldpc_b200.spacetime import detector_model_text, flatten_detector_model

DEM = """
# a comment line
error(0.01) D0 D1
error(0.02) D1 L0          # trailing comment
repeat 2 {
    error(0.03) D0 D2 ^ D2 L1
    shift_detectors(0, 1) 2
}
detector(1, 2) D3
error(0.25) L0
logical_observable L2
"""


def test_fields_of_a_model_with_loops_and_shifts():
    c = DetectorSpacetimeCode(DEM)
    assert c.num_detectors == 8 and c.num_faults == 5            # detector D3 after two shifts of 2 is D7
    want = np.zeros((8, 5), dtype=np.uint32)
    want[[0, 1], 0] = 1
    want[1, 1] = 1
    want[0, 2] = 1                                               # D2 ^ D2 cancels
    want[2, 3] = 1                                               # second pass: offset 2
    assert np.array_equal(c.fault_check_matrix.toarray(), want)
    assert c.fault_check_matrix.dtype == np.uint32 and c.fault_map.dtype == np.uint32
    assert np.array_equal(c.fault_map.toarray(), np.array([[0, 1, 0, 0, 1], [0, 0, 1, 1, 0], [0, 0, 0, 0, 0]], dtype=np.uint32))
    assert np.array_equal(c.fault_priors, [0.01, 0.02, 0.03, 0.03, 0.25])


def test_text_round_trip_and_wrapping():
    c = DetectorSpacetimeCode(DEM)
    again = DetectorSpacetimeCode(c.text())
    assert (again.fault_check_matrix != c.fault_check_matrix).nnz == 0 and (again.fault_map != c.fault_map).nnz == 0
    assert np.array_equal(again.fault_priors, c.fault_priors)
    assert again.fault_check_matrix.shape == c.fault_check_matrix.shape and again.fault_map.shape == c.fault_map.shape

    class Fields:                                                # circuits.DetectorErrorModel has these three fields
        fault_check_matrix, fault_map, fault_priors = c.fault_check_matrix, c.fault_map, c.fault_priors
    w = DetectorSpacetimeCode(Fields())
    assert (w.fault_check_matrix != c.fault_check_matrix).nnz == 0 and np.array_equal(w.fault_priors, c.fault_priors)

    class StimLike:                                              # stim.DetectorErrorModel: flattened() then str()
        def flattened(self):
            return self

        def __str__(self):
            return "error(0.5) D1 L0\nerror(0.125) D0\n"
    s = DetectorSpacetimeCode(StimLike())
    assert np.array_equal(s.fault_check_matrix.toarray(), [[0, 1], [1, 0]]) and np.array_equal(s.fault_map.toarray(), [[1, 0]])


def test_nested_repeat_and_empty_model():
    errors, max_d, max_l = flatten_detector_model("repeat 3 {\n repeat 2 {\n error(0.1) D0\n shift_detectors 1\n }\n}\n")
    assert [e[1] for e in errors] == [[k] for k in range(6)] and max_d == 5 and max_l == -1
    c = DetectorSpacetimeCode("")
    assert c.fault_check_matrix.shape == (0, 0) and c.fault_map.shape == (0, 0) and c.fault_priors.size == 0
    assert detector_model_text(c.fault_check_matrix, c.fault_map, c.fault_priors) == "\n"


@pytest.mark.parametrize("bad", ["error D0", "error(0.1, 0.2) D0", "error(0.1) D0 Q3", "repeat 2 {\nerror(0.1) D0\n",
                                 "frobnicate(1) D0"])
def test_malformed_models_raise(bad):
    with pytest.raises(ValueError):
        DetectorSpacetimeCode(bad)


def test_decoder_inputs_line_up():
    """What BPDetectorCorrect hands ldpc (``_experiment.py:137-140``): H = fault_check_matrix, one prior per column."""
    c = DetectorSpacetimeCode(DEM)
    assert c.fault_priors.shape == (c.fault_check_matrix.shape[1],) == (c.fault_map.shape[1],)
    assert c.fault_check_matrix.has_sorted_indices


@pytest.mark.parametrize("seed", range(8))
def test_random_models_survive_the_text_round_trip(seed):
    from scipy import sparse
    rng = np.random.default_rng(seed)
    nd, no, nf = int(rng.integers(1, 40)), int(rng.integers(0, 6)), int(rng.integers(1, 60))
    fc = sparse.csr_matrix((rng.random((nd, nf)) < 0.1).astype(np.uint32))
    fm = sparse.csr_matrix((rng.random((no, nf)) < 0.2).astype(np.uint32))
    pr = rng.uniform(1e-6, 0.4, nf)
    c = DetectorSpacetimeCode(detector_model_text(fc, fm, pr))
    assert c.fault_check_matrix.shape == (nd, nf) and c.fault_map.shape == (no, nf)
    assert (c.fault_check_matrix != fc).nnz == 0 and (c.fault_map != fm).nnz == 0
    assert np.array_equal(c.fault_priors, pr)                    # repr() of a double parses back to the same double

"""tcgen05 matcher: results must be IDENTICAL to the exact fp32 SIMT kernel and to the CPU oracle."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu():
    from imagestitch_b200 import gpu as g
    assert g.device_count() > 0
    yield g
    g.set_matcher("tc")


def _unit(rng, n, d):
    x = rng.standard_normal((n, d)).astype(np.float32)
    return x / np.linalg.norm(x, axis=1, keepdims=True)


@pytest.mark.parametrize("na,nb,d", [(300, 257, 128), (1000, 900, 64), (129, 2, 128), (4100, 3900, 128), (64, 5000, 32)])
def test_tc_equals_simt_and_oracle(gpu, na, nb, d):
    from oracle import surf
    rng = np.random.default_rng(na * 7 + nb)
    A, B = _unit(rng, na, d), _unit(rng, nb, d)
    # make a share of the queries true matches (small perturbations of train rows), like real overlapping tiles
    k = min(na, nb) // 3
    A[:k] = B[rng.permutation(nb)[:k]] + 0.05 * rng.standard_normal((k, d)).astype(np.float32)
    for ratio in (0.75, 0.99):
        gpu.set_matcher("tc")
        m_tc = gpu.match_descriptors(A, B, 2, ratio)
        fb = gpu.last_match_fallbacks()
        gpu.set_matcher("tc_1sm")
        m_cl = gpu.match_descriptors(A, B, 2, ratio)
        fb_cl = gpu.last_match_fallbacks()
        gpu.set_matcher("simt")
        m_simt = gpu.match_descriptors(A, B, 2, ratio)
        assert np.array_equal(m_cl, m_simt)
        assert fb_cl <= max(2, na // 50)
        m_or = surf.match_l2_ratio(A, B, ratio)
        assert np.array_equal(m_tc, m_simt)
        assert np.array_equal(m_tc, m_or)
        assert fb <= max(2, na // 50), "fallback scan used for %d of %d queries" % (fb, na)
        # every operand scheme: same matches, few exact rescans, and no rescored candidate outside the scheme's error bound
        for scheme, limit in (("tc_bf16x3", 50), ("tc_f16x2", 50), ("tc_f16x1", 20)):
            gpu.set_matcher(scheme)
            m_s = gpu.match_descriptors(A, B, 2, ratio)
            assert np.array_equal(m_s, m_simt), scheme
            assert gpu.last_match_bound_violations() == 0, scheme
            assert gpu.last_match_fallbacks() <= max(2, na // limit), (scheme, gpu.last_match_fallbacks())


def test_tc_duplicates_and_ties(gpu):
    """Exact duplicates in the train set: ties must resolve to the lower train index, ratio test then fails."""
    from oracle import surf
    rng = np.random.default_rng(5)
    A = _unit(rng, 200, 128)
    B = np.concatenate([A[:50], A[:50], _unit(rng, 300, 128)])
    gpu.set_matcher("tc")
    m_tc = gpu.match_descriptors(A, B, 2, 0.75)
    assert np.array_equal(m_tc, surf.match_l2_ratio(A, B, 0.75))


def test_tc_unnormalised_descriptors(gpu):
    """SIFT-like magnitudes (0..255): the norm terms folded into the GEMM and the error guard must still hold."""
    from oracle import surf
    rng = np.random.default_rng(9)
    A = np.abs(rng.standard_normal((700, 128))).astype(np.float32) * 60
    B = np.abs(rng.standard_normal((650, 128))).astype(np.float32) * 60
    A[:100] = B[:100] + rng.standard_normal((100, 128)).astype(np.float32)
    want = surf.match_l2_ratio(A, B, 0.8)
    gpu.set_matcher("tc_bf16x3")
    assert np.array_equal(gpu.match_descriptors(A, B, 2, 0.8), want)
    assert gpu.last_match_bound_violations() == 0 and gpu.last_match_fallbacks() <= len(A) // 10
    # featureType 1 (SIFT) always takes the bf16 operands, whatever the default scheme is
    gpu.set_matcher("tc")
    assert np.array_equal(gpu.match_descriptors(A, B, 1, 0.8), want)
    assert gpu.last_match_fallbacks() <= len(A) // 10
    # the fp16 schemes have no range for these norms: the pair is flagged and every query is rescanned exactly
    gpu.set_matcher("tc_f16x2")
    assert np.array_equal(gpu.match_descriptors(A, B, 2, 0.8), want)
    assert gpu.last_match_fallbacks() == len(A)


def test_tc_fp16_subnormal_operands(gpu):
    """Descriptors with many tiny components: the low half of the query split and the small train components are fp16
    subnormals.  The error bound assumes the tensor core does not flush them."""
    from oracle import surf
    rng = np.random.default_rng(21)
    A, B = _unit(rng, 1500, 128), _unit(rng, 1400, 128)
    scale = np.where(rng.random((1, 128)) < 0.7, 1e-4, 1.0).astype(np.float32)      # 70 % of the components ~1e-5
    A = (A * scale); B = (B * scale)
    A /= np.linalg.norm(A, axis=1, keepdims=True); B /= np.linalg.norm(B, axis=1, keepdims=True)
    A[:300] = B[:300] + 0.02 * rng.standard_normal((300, 128)).astype(np.float32) * scale
    want = surf.match_l2_ratio(A, B, 0.75)
    for scheme in ("tc_f16x2", "tc_f16x1", "tc_bf16x3"):
        gpu.set_matcher(scheme)
        assert np.array_equal(gpu.match_descriptors(A, B, 2, 0.75), want), scheme
        assert gpu.last_match_bound_violations() == 0, scheme


def test_tc_on_surf_descriptors(gpu, synth_pair_rois):
    from oracle import surf
    roiA, roiB, _ = synth_pair_rois
    kA, dA = gpu.surf_detect_and_describe(roiA)
    kB, dB = gpu.surf_detect_and_describe(roiB)
    gpu.set_matcher("tc")
    m = gpu.match_descriptors(dA, dB, 2, 0.75)
    assert np.array_equal(m, surf.match_l2_ratio(dA, dB, 0.75))
    assert gpu.last_match_fallbacks() <= len(dA) // 50
    assert gpu.last_match_bound_violations() == 0

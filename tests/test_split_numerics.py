"""Host-side check of the numerics argument behind csrc/flow_tc.cu: a float32 value splits EXACTLY into
three bfloat16 parts by truncation, and the six retained cross products of a 3 x BF16 split reproduce a float32 dot
product to ~1e-7 of its scale -- the budget that lets 16-bit tensor-core MMAs stand in for float32 FFMA GEMMs.  Pure NumPy
(no GPU): the device kernels are compared with the FFMA plan and the float64 oracle in tests/test_gpu_models.py."""
import numpy as np


def split3(a):
    """a = p1 + p2 + p3, each part a float32 whose low 16 bits are zero (= a bfloat16), as flow_tc.cu::split3."""
    a = np.asarray(a, np.float32)
    top = lambda v: (v.view(np.uint32) & np.uint32(0xFFFF0000)).view(np.float32)
    p1 = top(a)
    r1 = a - p1
    p2 = top(r1)
    r2 = r1 - p2
    return p1, p2, top(r2)


def test_split_is_exact_for_normal_float32():
    # (below 2^-110 the residuals go subnormal and are truncated away: irrelevant at the magnitudes of activations,
    # weights and gradients)
    rng = np.random.default_rng(0)
    a = np.concatenate([rng.normal(size=100000), rng.normal(size=1000) * 1e-20, rng.normal(size=1000) * 1e20,
                        [0.0, -0.0, 1.0, -1.0, 1.5e-33]]).astype(np.float32)  # zero and |a| > 2^-110 (residuals stay normal)
    p1, p2, p3 = split3(a)
    for p in (p1, p2, p3):
        assert np.all((p.view(np.uint32) & np.uint32(0xFFFF)) == 0)  # representable in bfloat16
    total = p1.astype(np.float64) + p2.astype(np.float64) + p3.astype(np.float64)
    assert np.array_equal(total.astype(np.float32), a)
    # truncation keeps 8 significant bits per part: the residual after three parts is below 2^-24 of |a|
    assert np.all(np.abs(total - a.astype(np.float64)) <= np.abs(a.astype(np.float64)) * 2.0**-23)


def test_six_products_of_the_split_match_float32_dot_products():
    """C = A B with A [64, 112], B [112, 96] (the shapes of one coupling-block tile): the six products a1b1, a1b2, a2b1,
    a1b3, a2b2, a3b1 (each exact in float32, accumulated here in float64 to isolate the split error) against the float64
    product of the float32 operands."""
    rng = np.random.default_rng(1)
    A = np.tanh(rng.normal(size=(64, 112))).astype(np.float32)          # hidden activations
    B = (rng.normal(size=(112, 96)) * 0.2).astype(np.float32)           # heads weights
    a, b = split3(A), split3(B)
    keep = [(0, 0), (0, 1), (1, 0), (0, 2), (1, 1), (2, 0)]
    C = sum(a[i].astype(np.float64) @ b[j].astype(np.float64) for i, j in keep)
    ref = A.astype(np.float64) @ B.astype(np.float64)
    scale = np.abs(A.astype(np.float64)) @ np.abs(B.astype(np.float64))
    err = np.abs(C - ref) / scale
    assert err.max() < 2.0**-22, err.max()   # dropped terms a2b3, a3b2, a3b3: 2^-24 .. 2^-32 of the scale
    # two parts only (the 2 x BF16 scheme) would NOT do: its error is ~2^-16 of the scale
    C2 = sum(a[i].astype(np.float64) @ b[j].astype(np.float64) for i, j in [(0, 0), (0, 1), (1, 0)])
    assert (np.abs(C2 - ref) / scale).max() > 2.0**-19

"""CPU tier: the arithmetic of the FP16-split decode (csrc/decode_tc16.cu) restated in numpy -- per-basis and per-frame
power-of-two scaling into [2^9, 2^10), hi = fp16(v), lo = fp16(v - hi), three products accumulated in fp32, exact
unscaling -- against fp64, for the reference's basis widths (K = 85 / 180, speech_anime/config/model/dgrad.py:75-92) and
for bases / coefficients with the magnitudes a trained model has (means of the scale part near 1 on the diagonal entries,
coefficients of very different size between frames).  The bound asserted is the one the GPU tests use for the kernel."""
import numpy as np

from deformation import workloads as W


def _pow2_scale(m):
    """The power of two that brings m into [2^9, 2^10) (k_split16 / tc16_build_basis)."""
    m = np.asarray(m, dtype=np.float64)
    _, q = np.frexp(np.where(m > 0, m, 1.0))
    return np.ldexp(1.0, 10 - q)


def _split16(v):
    hi = v.astype(np.float16)
    lo = (v - hi.astype(np.float64)).astype(np.float16)
    return hi.astype(np.float32), lo.astype(np.float32)


def fp16_split_decode(x, Wm, mean):
    """x [n, K] fp32, Wm [rows, K] fp32, mean [rows] -> [n, rows] fp32 the way the kernel computes it."""
    n, K = x.shape
    sw = float(_pow2_scale(max(np.abs(Wm).max(), np.abs(mean).max())))
    sx = _pow2_scale(np.maximum(np.abs(x).max(axis=1), 1.0))                       # per frame, the 1 = the means' column
    Wa = np.concatenate([Wm.astype(np.float64), mean.astype(np.float64)[:, None]], axis=1) * sw
    Xa = np.concatenate([x.astype(np.float64), np.ones((n, 1))], axis=1) * sx[:, None]
    wh, wl = _split16(Wa)
    xh, xl = _split16(Xa)
    acc = xh @ wh.T + xl @ wh.T + xh @ wl.T                                        # float32 matmuls = fp32 accumulation
    return (acc * (1.0 / (sx[:, None] * sw)).astype(np.float32)).astype(np.float32)


def test_fp16_split_matches_fp64_on_reference_widths():
    cs, ms, cr, mr = W.random_pca(800, seed=1)
    xs, xr = W.random_coeffs(96, seed=2)
    for x, Wm, mean in ((xs, cs, ms), (xr, cr, mr)):
        exact = x.astype(np.float64) @ Wm.astype(np.float64).T + mean
        got = fp16_split_decode(x, Wm, mean)
        fp32 = (x @ Wm.T + mean).astype(np.float64)
        assert np.abs(got - exact).max() <= 2e-7
        assert np.abs(got - exact).max() <= 4 * max(np.abs(fp32 - exact).max(), 2e-8)      # as good as fp32 FMA, within a small factor


def test_fp16_split_with_trained_model_magnitudes():
    rng = np.random.default_rng(7)
    rows, K, n = 1800, 85, 64
    q, _ = np.linalg.qr(rng.standard_normal((rows, K)))                            # PCA components: orthonormal columns, |w| < 1
    Wm = q.astype(np.float32)
    mean = (0.01 * rng.standard_normal(rows)).astype(np.float32)
    mean[::6] += 1.0                                                               # s00-like entries: identity on the diagonal
    x = rng.standard_normal((n, K)).astype(np.float32) * np.geomspace(30.0, 0.01, K).astype(np.float32)   # decaying spectrum
    x[3] *= 1e-5                                                                   # a frame of tiny coefficients
    x[5] *= 40.0                                                                   # and one of huge ones
    exact = x.astype(np.float64) @ Wm.astype(np.float64).T + mean
    got = fp16_split_decode(x, Wm, mean).astype(np.float64)
    fp32 = (x @ Wm.T + mean).astype(np.float64)                                    # what torch F.linear computes (fp32 FMA)
    scale = np.maximum(np.abs(exact).max(axis=1), 1.0)                             # per frame: error relative to its own magnitude
    e16, e32 = np.abs(got - exact).max(axis=1) / scale, np.abs(fp32 - exact).max(axis=1) / scale
    assert e16.max() <= 1.5e-6 and e16.max() <= 2 * e32.max()                      # at these magnitudes fp32 rounding itself is 7e-7
    assert np.abs(got[3] - exact[3]).max() <= 2e-7                                 # the tiny frame keeps absolute accuracy (means dominate)

"""CPU: an executable model of the arithmetic of k2_peaks_fast (csrc/opp_kernels.cu row_taps / col_phase) against the
oracle's cv::GaussianBlur restatement on the replicated image.

At an integer scale S the up-sampled map is S x S blocks of one feature value.  For a Gaussian radius R <= S the fast
kernel therefore (1) runs the row pass on FEATURE rows only, each output pixel seeing at most three distinct neighbours
(a, b, c) - five (a2, a, b, c, c2) for S < R <= 2S - plus the single taps at distance exactly S and 2S that
BORDER_REFLECT_101 sends to another neighbour at the image edge (the *s operands) - and (2) runs the column pass per
feature row with the same neighbour structure.  Every product and sum is
the IEEE operation cv::GaussianBlur performs on the same operands in the same order, so the result must be bit-identical
to blurring the materialised image.  This test restates that decomposition in numpy float32 and checks exactly that, for
both scales the kernel is compiled for and every radius, including the symmetric-small row forms (k = 3, 5)."""
import numpy as np
import pytest

from oracle.oracle import Oracle

f32 = np.float32


def make_win(m2, m1, z, p1, p2, i, n, zero):
    """The nine operands (a2s, a2, as, a, b, c, cs, c2, c2s) the kernel's make_win builds for cell i of n: the cell, its
    two neighbours on either side, and the values the border rule substitutes for the single taps at distance exactly S
    and 2S (the *s operands)."""
    z0 = f32(0)
    if zero:
        a = m1 if i > 0 else z0
        a2 = m2 if i > 1 else z0
        c = p1 if i < n - 1 else z0
        c2 = p2 if i < n - 2 else z0
        return (a2, a2, a, a, z, c, c, c2, c2)
    a = m1 if i > 0 else z
    a_s = m1 if i > 0 else p1
    a2 = m2 if i > 1 else (m1 if i == 1 else p1)
    a2s = m2 if i > 1 else (z if i == 1 else p2)
    c = p1 if i < n - 1 else z
    c_s = p1 if i < n - 1 else m1
    c2 = p2 if i < n - 2 else (p1 if i == n - 2 else m1)
    c2s = p2 if i < n - 2 else (z if i == n - 2 else m2)
    return (a2s, a2, a_s, a, z, c, c_s, c2, c2s)


def win_src(win, S, o):
    """value of the replicated line at offset o from the start of the cell (o in [-2S, 3S-1])"""
    a2s, a2, a_s, a, b, c, c_s, c2, c2s = win
    if 0 <= o < S:
        return b
    if o < 0:
        d = -o
        return a if d < S else a_s if d == S else a2 if d < 2 * S else a2s
    d = o - S + 1
    return c if d < S else c_s if d == S else c2 if d < 2 * S else c2s


def row_taps(k, S, R, win, small=True):
    src = lambda o: win_src(win, S, o)  # noqa: E731
    out = []
    for q in range(S):
        if R == 1 and small:
            s = f32(f32(src(q) * k[1]) + f32(f32(src(q - 1) + src(q + 1)) * k[2]))
        elif R == 2 and small:
            s = f32(f32(src(q) * k[2]) + f32(f32(src(q - 1) + src(q + 1)) * k[3]))
            s = f32(s + f32(f32(src(q - 2) + src(q + 2)) * k[4]))
        else:
            s = None
            for j in range(2 * R + 1):
                pr = f32(k[j] * src(q + j - R))
                s = pr if s is None else f32(s + pr)
        out.append(s)
    return out


def col_phase(k, S, R, ph, win):
    s = f32(k[R] * win[4])
    for j in range(1, R + 1):
        s = f32(s + f32(k[R + j] * f32(win_src(win, S, ph + j) + win_src(win, S, ph - j))))
    return s


def fast_model(F, S, ksize, zero=False, taps=None):
    h, w = F.shape
    R = ksize // 2
    k = taps if taps is not None else (Oracle.gauss_kernel(ksize) if ksize > 1 else np.ones(1, np.float32))
    at = lambda A, r, c: A[min(max(r, 0), A.shape[0] - 1), min(max(c, 0), A.shape[1] - 1)]  # noqa: E731  (clamped: the value is unused when out of range)
    rrow = np.zeros((h, S * w), np.float32)  # row pass, one row per FEATURE row
    for r in range(h):
        for c in range(w):
            win = make_win(at(F, r, c - 2), at(F, r, c - 1), F[r, c], at(F, r, c + 1), at(F, r, c + 2), c, w, zero)
            rrow[r, S * c:S * c + S] = [F[r, c]] * S if ksize == 1 else row_taps(k, S, R, win, small=not zero)
    out = np.zeros((S * h, S * w), np.float32)
    for i in range(h):
        for x in range(S * w):
            win = make_win(at(rrow, i - 2, x), at(rrow, i - 1, x), rrow[i, x], at(rrow, i + 1, x), at(rrow, i + 2, x), i, h, zero)
            for ph in range(S):
                out[S * i + ph, x] = rrow[i, x] if ksize == 1 else col_phase(k, S, R, ph, win)
    return out


@pytest.mark.parametrize("S,ksizes", [(8, (1, 3, 5, 7, 9, 11, 13, 15, 17, 19, 21, 23, 25, 27, 29, 31, 33)), (4, (1, 3, 5, 7, 9, 11, 13, 15, 17))])
def test_feature_level_arithmetic_equals_blurring_the_replicated_image(S, ksizes):
    """R <= S: three distinct neighbours; S < R <= 2S (k = 19..33 at x8, the Python graph's 25 included): five."""
    rng = np.random.default_rng(S)
    for ksize in ksizes:
        for (h, w) in ((5, 6), (2, 7), (3, 2), (3, 3), (4, 9)):
            if ksize // 2 > S and min(h, w) < 3:
                continue  # the kernel takes R > S only for maps of at least 3 x 3 cells
            F = rng.random((h, w), dtype=np.float32)
            img = np.repeat(np.repeat(F, S, 0), S, 1)
            want = Oracle.gauss_blur(img, ksize)
            got = fast_model(F, S, ksize)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (S, ksize, h, w)


@pytest.mark.parametrize("S,ksize", [(8, 25), (8, 17), (8, 33), (4, 13), (4, 5), (8, 3)])
def test_feature_level_arithmetic_with_zero_border(S, ksize):
    """The Python graph's 'SAME' zero padding (post_process.py:25-26) through the same operand scheme: cells beyond the
    border stand as 0 and every tap runs the general left-to-right row form."""
    rng = np.random.default_rng(100 + ksize)
    taps = Oracle.cdf_kernel(ksize)
    for (h, w) in ((5, 6), (3, 3), (4, 9)):
        F = rng.random((h, w), dtype=np.float32)
        img = np.repeat(np.repeat(F, S, 0), S, 1)
        want = Oracle.smooth_zero_pad(img, taps)
        got = fast_model(F, S, ksize, zero=True, taps=taps)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (S, ksize, h, w)


def reflect101(p, n):
    if p < 0:
        return -p
    if p >= n:
        return 2 * (n - 1) - p
    return p


def generic_rep_model(F, S, ksize, zero, taps, GTX=64, GTY=32):
    """k2_peaks_generic_rep, tile by tile, with the kernel's own index formulas (f_lo, nfr, row map with an all-zero
    row, column expansion rx / S): returns the smoothed image."""
    h, w = F.shape
    H, W, R = S * h, S * w, ksize // 2
    IW, IH, TW = GTX + 2 + 2 * R, GTY + 2 + 2 * R, GTX + 2
    out = np.full((H, W), np.nan, np.float32)
    clip = lambda v, n: min(max(v, 0), n - 1)  # noqa: E731
    for y0 in range(0, H, GTY):
        for x0 in range(0, W, GTX):
            ylo, yhi = max(y0 - 1 - R, 0), min(y0 + GTY + R, H - 1)
            f_lo = ylo // S
            nfr = yhi // S - f_lo + 1
            rowmap = []
            for i in range(IH):
                yy = y0 - 1 - R + i
                if zero:
                    m = nfr if (yy < 0 or yy >= H) else yy // S - f_lo
                else:
                    m = min(max(clip(reflect101(yy, H), H) // S - f_lo, 0), nfr - 1)
                rowmap.append(m)
            inn = np.zeros((nfr, IW), np.float32)
            for fr in range(nfr):
                for t in range(IW):
                    xx = x0 - 1 - R + t
                    inn[fr, t] = f32(0) if (zero and (xx < 0 or xx >= W)) else F[f_lo + fr, clip(reflect101(xx, W), W) // S]
            tmp = np.zeros((nfr + 1, TW), np.float32)
            for fr in range(nfr):
                for c in range(TW):
                    q = inn[fr, c:c + ksize]
                    if ksize == 3 and not zero:
                        s = f32(f32(q[R] * taps[R]) + f32(f32(q[R - 1] + q[R + 1]) * taps[R + 1]))
                    elif ksize == 5 and not zero:
                        s = f32(f32(q[R] * taps[R]) + f32(f32(q[R - 1] + q[R + 1]) * taps[R + 1]))
                        s = f32(s + f32(f32(q[R - 2] + q[R + 2]) * taps[R + 2]))
                    else:
                        s = f32(taps[0] * q[0])
                        for j in range(1, ksize):
                            s = f32(s + f32(taps[j] * q[j]))
                    tmp[fr, c] = s
            for r in range(1, GTY + 1):
                y = y0 - 1 + r
                if y >= H:
                    break
                for c in range(1, TW - 1):
                    x = x0 - 1 + c
                    if x >= W:
                        break
                    rm = lambda j: rowmap[r + R + j]  # noqa: E731
                    s = f32(taps[R] * tmp[rm(0), c])
                    for j in range(1, R + 1):
                        s = f32(s + f32(taps[R + j] * f32(tmp[rm(j), c] + tmp[rm(-j), c])))
                    out[y, x] = s
    return out


@pytest.mark.parametrize("zero", [False, True])
def test_replication_aware_generic_kernel_model(zero):
    """The row map of k2_peaks_generic_rep (image row -> feature row of the row-pass result, REFLECT_101 or an all-zero row
    for tf 'SAME' padding) gives the same smoothed image as filtering the materialised map, bit for bit."""
    rng = np.random.default_rng(5)
    for (S, ksize, h, w) in ((8, 25, 5, 9), (8, 19, 9, 3), (4, 13, 11, 17), (2, 9, 20, 37), (1, 7, 40, 70), (8, 3, 5, 9), (4, 5, 9, 17)):
        F = rng.random((h, w), dtype=np.float32)
        img = np.repeat(np.repeat(F, S, 0), S, 1)
        if zero:
            taps = Oracle.cdf_kernel(ksize)
            want = Oracle.smooth_zero_pad(img, taps)
        else:
            taps = Oracle.gauss_kernel(ksize)
            want = Oracle.gauss_blur(img, ksize)
        got = generic_rep_model(F, S, ksize, zero, taps)
        assert not np.isnan(got).any()
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (S, ksize, h, w, zero)

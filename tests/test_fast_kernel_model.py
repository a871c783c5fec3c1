"""CPU: an executable model of the arithmetic of k2_peaks_fast (csrc/opp_kernels.cu row_taps / col_phase) against the
oracle's cv::GaussianBlur restatement on the replicated image.

At an integer scale S the up-sampled map is S x S blocks of one feature value.  For a Gaussian radius R <= S the fast
kernel therefore (1) runs the row pass on FEATURE rows only, each output pixel seeing at most three distinct neighbours
(a, b, c) - plus, when R == S, one tap that BORDER_REFLECT_101 sends to the other neighbour at the image edge (a_sp /
c_sp) - and (2) runs the column pass per feature row with the same three-neighbour structure.  Every product and sum is
the IEEE operation cv::GaussianBlur performs on the same operands in the same order, so the result must be bit-identical
to blurring the materialised image.  This test restates that decomposition in numpy float32 and checks exactly that, for
both scales the kernel is compiled for and every radius, including the symmetric-small row forms (k = 3, 5)."""
import numpy as np
import pytest

from oracle.oracle import Oracle

f32 = np.float32


def row_taps(k, S, R, a, b, c, a_sp, c_sp):
    def src(o):
        if o < 0:
            return a_sp if (R == S and o == -S) else a
        if o < S:
            return b
        return c_sp if (R == S and o == 2 * S - 1) else c
    out = []
    for q in range(S):
        if R == 1:
            s = f32(f32(src(q) * k[1]) + f32(f32(src(q - 1) + src(q + 1)) * k[2]))
        elif R == 2:
            s = f32(f32(src(q) * k[2]) + f32(f32(src(q - 1) + src(q + 1)) * k[3]))
            s = f32(s + f32(f32(src(q - 2) + src(q + 2)) * k[4]))
        else:
            s = None
            for j in range(2 * R + 1):
                pr = f32(k[j] * src(q + j - R))
                s = pr if s is None else f32(s + pr)
        out.append(s)
    return out


def col_phase(k, S, R, ph, a, b, c, a_sp, c_sp):
    s = f32(k[R] * b)
    for j in range(1, R + 1):
        up = b if ph + j < S else (c_sp if (R == S and ph == S - 1 and j == R) else c)
        dn = b if ph - j >= 0 else (a_sp if (R == S and ph == 0 and j == R) else a)
        s = f32(s + f32(k[R + j] * f32(up + dn)))
    return s


def fast_model(F, S, ksize):
    h, w = F.shape
    R = ksize // 2
    k = Oracle.gauss_kernel(ksize) if ksize > 1 else np.ones(1, np.float32)
    rrow = np.zeros((h, S * w), np.float32)  # row pass, one row per FEATURE row
    for r in range(h):
        for c in range(w):
            b = F[r, c]
            a_raw, c_raw = F[r, max(c - 1, 0)], F[r, min(c + 1, w - 1)]
            a, cc = (a_raw if c > 0 else b), (c_raw if c < w - 1 else b)
            a_sp, c_sp = (a_raw if c > 0 else c_raw), (c_raw if c < w - 1 else a_raw)
            rrow[r, S * c:S * c + S] = [b] * S if ksize == 1 else row_taps(k, S, R, a, b, cc, a_sp, c_sp)
    out = np.zeros((S * h, S * w), np.float32)
    for i in range(h):
        top, bot = i == 0, i == h - 1
        for x in range(S * w):
            b = rrow[i, x]
            a = rrow[i - 1, x] if i > 0 else f32(0)
            c = rrow[i + 1, x] if i < h - 1 else f32(0)
            for ph in range(S):
                out[S * i + ph, x] = b if ksize == 1 else col_phase(k, S, R, ph, b if top else a, b, b if bot else c, c if top else a, a if bot else c)
    return out


@pytest.mark.parametrize("S,ksizes", [(8, (1, 3, 5, 7, 9, 11, 13, 15, 17)), (4, (1, 3, 5, 7, 9))])
def test_feature_level_arithmetic_equals_blurring_the_replicated_image(S, ksizes):
    rng = np.random.default_rng(S)
    for ksize in ksizes:
        for (h, w) in ((5, 6), (2, 7), (3, 2)):
            F = rng.random((h, w), dtype=np.float32)
            img = np.repeat(np.repeat(F, S, 0), S, 1)
            want = Oracle.gauss_blur(img, ksize)
            got = fast_model(F, S, ksize)
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

"""CPU: the oracle (plain-C restatement) against the committed golden vectors minted from the
reference's own build and cv2 (scripts/make_golden.py), and against the live reference build / cv2
where they are present in the container."""
import glob
import os

import numpy as np
import pytest

from oracle.oracle import Oracle, Reference, CAND_DT, ref_available
import helpers

GOLD = os.path.join(os.path.dirname(__file__), "golden")
FRAMES = sorted(glob.glob(os.path.join(GOLD, "frame_*.npz")))


def test_golden_present():
    assert len(FRAMES) >= 8 and os.path.exists(os.path.join(GOLD, "opencv_pieces.npz"))


@pytest.mark.parametrize("path", FRAMES, ids=[os.path.basename(p)[:-4] for p in FRAMES])
def test_oracle_reproduces_reference_humans(path):
    g = np.load(path)
    h, w, H, W, k = [int(v) for v in g["geom"]]
    o = Oracle(h, w, H, W, k).run(g["conf"], g["paf"])
    assert helpers.humans_equal(o["humans"], g["humans_ref"]) is None
    assert np.array_equal(o["peaks"], g["peaks"])
    for p in range(19):
        assert np.array_equal(o["conns"][p], g["conns_%02d" % p])
    assert np.array_equal(o["hrefs"]["parts"], g["hrefs"]["parts"])
    assert [o["n_incomplete"], o["n_merges"], o["flags"]] == g["counts"].tolist()
    # the lazy mode (PAF sampled on demand instead of materialised) is the same arithmetic
    ol = Oracle(h, w, H, W, k).run(g["conf"], g["paf"], lazy=True)
    assert helpers.humans_equal(ol["humans"], g["humans_ref"]) is None


def test_opencv_pieces_bit_exact():
    g = np.load(os.path.join(GOLD, "opencv_pieces.npz"))
    for k in range(1, 64, 2):
        assert np.array_equal(Oracle.gauss_kernel(k), g["taps_%d" % k]), k
    src = g["resize_src"]
    for key in g.files:
        if key.startswith("resize_") and key != "resize_src":
            H, W = [int(v) for v in key[7:].split("x")]
            assert np.array_equal(Oracle.resize_area(src, H, W), g[key]), key
    for k in (1, 3, 5, 7, 9, 13, 17, 25, 31):
        assert np.array_equal(Oracle.gauss_blur(g["blur_src"], k), g["blur_%d" % k]), k
        assert np.array_equal(Oracle.gauss_blur(g["blur_src2"], k), g["blur2_%d" % k]), k
    assert np.array_equal(Oracle.max_pool(g["blur_src2"]), g["dilate"])


def test_x8_upsample_is_replication():
    rng = np.random.default_rng(3)
    a = rng.random((46, 54), dtype=np.float32)
    assert np.array_equal(Oracle.resize_area(a, 368, 432), np.repeat(np.repeat(a, 8, 0), 8, 1))


def test_against_live_cv2():
    cv2 = pytest.importorskip("cv2")
    cv2.ipp.setUseIPP(False)
    cv2.setUseOptimized(False)
    cv2.setNumThreads(1)
    rng = np.random.default_rng(5)
    a = (rng.random((31, 29), dtype=np.float32) * 2 - 1).astype(np.float32)
    for (H, W) in [(248, 232), (100, 97), (31, 58), (63, 30)]:
        assert np.array_equal(Oracle.resize_area(a, H, W), cv2.resize(a, (W, H), interpolation=cv2.INTER_AREA))
    b = rng.random((64, 80), dtype=np.float32)
    for k in (3, 5, 9, 17, 21):
        assert np.array_equal(Oracle.gauss_blur(b, k), cv2.GaussianBlur(b, (k, k), 3.0))


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_std_sort_emulation_matches_libstdcxx():
    rng = np.random.default_rng(1)
    for n in list(range(0, 40)) + [100, 257, 1000, 5000]:
        for trial in range(4):
            v = np.zeros(n, CAND_DT)
            v["idx1"] = np.arange(n)
            v["idx2"] = rng.integers(0, 50, n)
            nd = [max(1, n), max(1, n // 4), 3, 1][trial]
            v["score"] = rng.integers(0, nd, n).astype(np.float32) / 7
            assert np.array_equal(Oracle.std_sort_desc(v), Reference.std_sort_desc(v)), (n, trial)
    for n in (1000, 4096):  # patterns that push introsort towards its depth limit
        for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]])):
            v = np.zeros(n, CAND_DT)
            v["idx1"] = np.arange(n)
            v["score"] = arr.astype(np.float32)
            assert np.array_equal(Oracle.std_sort_desc(v), Reference.std_sort_desc(v))


@pytest.mark.skipif(not ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_matches_live_reference_on_fresh_frames():
    from openpose_plus_b200 import synth
    orc, ref = Oracle(46, 54, 368, 432, 17), Reference(46, 54, 368, 432, 17)
    for seed, people, kw in [(900, 4, {}), (901, 33, {}), (902, 36, {"drop_limbs": (12,)}), (903, 6, {"noise": 1e-3})]:
        conf, paf = synth.render_frame(seed, people, **kw)
        o = orc.run(conf, paf)
        assert helpers.humans_equal(o["humans"], ref.run(conf, paf)) is None, seed


def test_python_path_kernel_and_zero_padded_smoothing():
    """Oracle variant 1 (the reference's Python graph, post_process.py:13-26): the separable taps reproduce
    _gauss_kernel's 2-D filter and the float32 separable zero-padded smoothing stays within float tolerance of a
    float64 2-D convolution (TensorFlow fixes no summation order: tolerance, not bits)."""
    st = pytest.importorskip("scipy.stats")
    sig = pytest.importorskip("scipy.signal")

    def gauss_kernel(ksize, nsig):                  # post_process.py:13-17, restated
        interval = (2 * nsig + 1.) / ksize
        x = np.linspace(-nsig - interval / 2., nsig + interval / 2., ksize + 1)
        y = np.diff(st.norm.cdf(x))
        k = np.sqrt(np.outer(y, y))
        return k / k.sum()

    for k in (25, 17, 9, 3, 1):
        t = Oracle.cdf_kernel(k).astype(np.float64)
        g2 = gauss_kernel(k, 3.0)
        assert np.abs(np.outer(t, t) - g2).max() <= 2e-7 * g2.max()
        assert abs(t.sum() - 1.0) < 1e-6 and (t > 0).all()
    rng = np.random.default_rng(0)
    img = rng.random((97, 131)).astype(np.float32)
    got = Oracle.smooth_zero_pad(img, Oracle.cdf_kernel(25))
    want = sig.convolve2d(img.astype(np.float64), gauss_kernel(25, 3.0), mode="same", boundary="fill")
    assert np.abs(got - want).max() < 1e-6
    # zero padding: a constant image fades towards the border (REFLECT_101 would keep it constant)
    flat = Oracle.smooth_zero_pad(np.ones((64, 64), np.float32), Oracle.cdf_kernel(25))
    assert abs(flat[32, 32] - 1) < 1e-6 and flat[0, 0] < 0.4 and flat[0, 32] < 0.7


def test_python_path_grouping_indexes_by_position():
    """Variant 1 groups like pafprocess (position in the vector); variant 0 like src/paf.cpp (stored id, stale after an
    erase).  They agree until the first merge and may differ after it; variant 1 never raises the stale-index flags."""
    from openpose_plus_b200 import synth
    c, p = synth.render_frame(201, n_people=32, drop_limbs=(12,))
    a, b = Oracle(46, 54, 368, 432, 17, variant=0).run(c, p), Oracle(46, 54, 368, 432, 17, variant=1).run(c, p)
    assert a["n_merges"] > 0 and b["n_merges"] > 0
    assert b["flags"] == 0
    c, p = synth.render_frame(3, n_people=5)
    o0, o1 = Oracle(46, 54, 368, 432, 17), Oracle(46, 54, 368, 432, 17)
    o1.L.orc_set_variant(o1.ctx, 0)
    assert o0.run(c, p)["n_humans"] == o1.run(c, p)["n_humans"]


def test_python_variant_oracle_reproduces_its_vectors():
    """Regression vectors of oracle variant 1 (scripts/make_golden_python_variant.py; unpinned, see its header)."""
    paths = sorted(glob.glob(os.path.join(GOLD, "pyvariant_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        g = np.load(path)
        h, w, oh, ow, k = [int(v) for v in g["geom"]]
        o = Oracle(h, w, oh, ow, k, variant=1).run(g["conf"], g["paf"])
        assert np.array_equal(o["peaks"].view(np.uint8), g["peaks"].view(np.uint8))
        assert np.array_equal(o["humans"].view(np.uint8), g["humans"].view(np.uint8))
        assert [o["n_incomplete"], o["n_merges"], o["flags"]] == g["counts"].tolist()


def test_block_skipping_bound_holds_for_every_kernel_size():
    """The peak kernels skip a block when every feature value its pixels depend on is <= t = 0.05f * (1 - 2^-13)
    (DESIGN.md section 4, K2 step 2).  Every operation of the filter is monotone in its inputs (positive taps, IEEE
    round-to-nearest), so the largest smoothed value such a block can reach is the one of the constant image t - which
    must not exceed THRESH_HEAT.  Checked here in the filter's own float32 arithmetic for every kernel size and both
    border rules."""
    thr = np.float32(0.05)
    t = np.float32(thr * np.float32(1.0 - 1.0 / 8192.0))
    img = np.full((80, 80), t, np.float32)
    for k in range(1, 64, 2):
        s = Oracle.gauss_blur(img, k)
        assert not (s > thr).any(), k
        z = Oracle.smooth_zero_pad(img, Oracle.cdf_kernel(k))
        assert not (z > thr).any(), k
    # and the bound is not vacuous: a constant image just above the threshold does produce values above it
    assert (Oracle.gauss_blur(np.full((80, 80), np.float32(0.0501), np.float32), 17) > thr).all()


def test_fp32_forms_used_by_the_limb_kernel_equal_the_references_double_forms():
    """csrc/opp_kernels.cu scores limbs without FP64 where an exact FP32 form exists; the three identities it relies on,
    checked on the CPU in IEEE arithmetic:
      (a) (float)sqrt((double)l2) == sqrtf((float)l2) for integer l2 < 2^24             (src/paf.cpp:91)
      (b) (int)((double)v + 0.5) == t + (v - t >= 0.5f), t = (int)v, for 0 <= v < 2^23  (roundpaf, src/paf.cpp:337)
      (c) min(0.0, 0.5 * H / (double)norm - 1.0) == 0 whenever norm <= 0.5f * H        (src/paf.cpp:115-116)"""
    d = np.arange(0, 2049, dtype=np.int64)
    l2 = (d[:, None] ** 2 + d[None, :] ** 2).ravel()
    l2 = np.unique(l2[l2 < (1 << 24)])
    assert np.array_equal(np.sqrt(l2.astype(np.float64)).astype(np.float32), np.sqrt(l2.astype(np.float32)))
    rng = np.random.default_rng(0)
    v = np.concatenate([rng.uniform(0, 4096, 2_000_000), np.arange(0, 4096, 0.5), np.nextafter(np.arange(0.5, 4096, 1.0), 0),
                        np.nextafter(np.arange(0.5, 4096, 1.0), 1e9)]).astype(np.float32)
    want = (v.astype(np.float64) + 0.5).astype(np.int64)
    t = v.astype(np.int32)
    got = t + ((v - t.astype(np.float32)) >= np.float32(0.5))
    assert np.array_equal(got, want)
    for H in (368, 736, 300, 97):
        norm = np.sqrt(l2[(l2 > 0) & (l2 < 4 * H * H)].astype(np.float32))
        inside = norm <= np.float32(0.5) * np.float32(H)
        pen = np.minimum(0.0, 0.5 * H / norm[inside].astype(np.float64) - 1.0)
        assert (pen == 0.0).all()


def test_pair_prefilter_forms_of_the_limb_kernel():
    """The limb kernel's first filter over (a, b) peak pairs (csrc/opp_kernels.cu pair_may_pass) rests on two facts,
    checked here in IEEE float32:
      (a) round_paf_small: for 0 <= v < 2^22, with r = v + 2^23 (ties to even) and d = v - (r - 2^23),
          (int)((double)v + 0.5) == (bits(r) - 0x4B000000) + (d >= 0.5f)                       (roundpaf, src/paf.cpp:337)
      (b) |v.x|, |v.y| <= 1 + 2^-23 for v = (dx / norm, dy / norm), norm = (float)sqrt((double)(dx^2 + dy^2)), so a sample
          in a cell with |px| + |py| <= thr (1 - 2^-13) scores fl(fl(vx px) + fl(vy py)) <= thr: it can never count
          towards `cnt` (src/paf.cpp:108-113), whatever the direction."""
    f32 = np.float32
    rng = np.random.default_rng(1)
    v = np.concatenate([rng.uniform(0, 40000, 2_000_000), np.arange(0, 40000, 0.5), np.nextafter(np.arange(0.5, 40000, 1.0), 0),
                        np.nextafter(np.arange(0.5, 40000, 1.0), 1e9), [0.0, 0.49999997, 0.5, 4194303.5, 4194303.0]]).astype(f32)
    want = (v.astype(np.float64) + 0.5).astype(np.int64)
    r = (v + f32(8388608.0)).astype(f32)
    d = (v - (r - f32(8388608.0)).astype(f32)).astype(f32)
    got = (r.view(np.int32).astype(np.int64) - 0x4B000000) + (d >= f32(0.5))
    assert np.array_equal(got, want)
    dd = np.arange(-1200, 1201, dtype=np.int64)
    dx, dy = np.meshgrid(dd, dd)
    dx, dy = dx.ravel(), dy.ravel()
    keep = (dx != 0) | (dy != 0)
    dx, dy = dx[keep], dy[keep]
    norm = np.sqrt((dx * dx + dy * dy).astype(np.float64)).astype(f32)
    vx, vy = (dx.astype(f32) / norm).astype(f32), (dy.astype(f32) / norm).astype(f32)
    assert max(np.abs(vx).max(), np.abs(vy).max()) <= 1.0 + 2.0 ** -23
    thr = f32(0.05)
    tw = f32(thr * f32(1.0 - 1.0 / 8192.0))
    # adversarial cell values on the bound: all of the magnitude in the axis the direction favours, and split evenly
    idx = rng.integers(0, len(vx), 400_000)
    for (px, py) in ((tw * np.sign(vx[idx]), 0 * vy[idx]), (0 * vx[idx], tw * np.sign(vy[idx])), (tw / 2 * np.sign(vx[idx]), tw / 2 * np.sign(vy[idx]))):
        px, py = px.astype(f32), py.astype(f32)
        assert ((np.abs(px) + np.abs(py)).astype(f32) <= tw).all()
        score = ((vx[idx] * px).astype(f32) + (vy[idx] * py).astype(f32)).astype(f32)
        assert not (score > thr).any()

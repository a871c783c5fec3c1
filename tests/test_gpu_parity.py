"""-m gpu: the CUDA path, called through the C-ABI, against the CPU oracle, stage by stage.
Integer outputs (peak coordinates and ids, limb assignments, person->part ids) and all scores are
required bit-exact (north_star allows 1e-5 relative on scores; the kernels do better)."""
import numpy as np
import pytest

from openpose_plus_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from openpose_plus_b200.engine import Engine
    from oracle.oracle import Oracle
    import helpers
    return Engine, Oracle, helpers


@pytest.mark.parametrize("ksize", [17, 13, 9])
def test_typical_frames_368x432(mods, ksize):
    Engine, Oracle, H = mods
    conf, paf = synth.render_batch(8, n_people=5, seed0=100)
    eng, orc = Engine(46, 54, gauss_kernel_size=ksize, max_batch=8), Oracle(46, 54, 368, 432, ksize)
    H.run_and_check(eng, orc, conf, paf, "k=%d" % ksize)


def test_crowded_frames(mods):
    Engine, Oracle, H = mods
    fr = [synth.render_frame(200 + i, n_people=30 + 2 * i, drop_limbs=(12,) if i % 2 else ()) for i in range(6)]
    conf, paf = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])
    eng, orc = Engine(46, 54, max_batch=6, max_humans=256), Oracle(46, 54, 368, 432, 17)
    H.run_and_check(eng, orc, conf, paf, "crowded")


def test_highres_736x864(mods):
    Engine, Oracle, H = mods
    conf, paf = synth.render_batch(3, n_people=12, feat_h=92, feat_w=108, seed0=300)
    eng, orc = Engine(92, 108, max_batch=4), Oracle(92, 108, 736, 864, 17)
    H.run_and_check(eng, orc, conf, paf, "736x864")


def test_single_frame_and_empty(mods):
    Engine, Oracle, H = mods
    eng, orc = Engine(46, 54, max_batch=2), Oracle(46, 54, 368, 432, 17)
    conf, paf = synth.render_batch(1, n_people=3, seed0=7)
    H.run_and_check(eng, orc, conf, paf, "single")
    z = np.zeros_like(conf), np.zeros_like(paf)
    humans, counts, flags = H.run_and_check(eng, orc, z[0], z[1], "empty maps")
    assert counts[0] == 0


def test_generic_kernel_sizes_and_scales(mods):
    Engine, Oracle, H = mods
    conf, paf = synth.render_batch(2, n_people=6, seed0=400)
    # small kernels on x8-replicated maps are all plateaus (thousands of tied peaks), so k = 1, 3, 5
    # (OpenCV's copy / symmetric-small row forms) are exercised at scale 1 and 2 instead
    for (oh, ow, k) in [(368, 432, 35), (368, 432, 41), (46 * 4, 54 * 4, 19), (46, 54, 1), (46, 54, 3), (92, 108, 5), (46, 108, 7), (300, 400, 17),
                        (369, 433, 9), (46 * 3, 54 * 3, 7), (46 * 4, 54 * 4, 9), (368, 432 * 2, 17)]:
        eng, orc = Engine(46, 54, oh, ow, gauss_kernel_size=k, max_batch=2, max_peaks_per_part=512), Oracle(46, 54, oh, ow, k)
        H.run_and_check(eng, orc, conf, paf, "generic %dx%d k=%d" % (oh, ow, k))


def test_non_integer_scales_from_feature_maps_and_from_the_materialised_map(mods, monkeypatch):
    """Non-integer scales: the generic peak kernel builds its tiles from the feature maps (horizontal 2-tap pass on the
    feature rows, vertical blend per image row = cv::resize's own operations) or, with OPP_GENERIC_VIA_MAP=1, reads a
    materialised map: both bit-exact against the oracle.  Large kernels, both border rules, odd sizes, noisy maps that
    reach every border, and scales close to 1 (where nearly every feature row is its own image row)."""
    Engine, Oracle, H = mods
    from openpose_plus_b200 import _capi as capi
    rng = np.random.default_rng(21)
    conf, paf = synth.render_batch(2, n_people=6, seed0=930)
    conf[:, :18] = np.maximum(conf[:, :18], (0.5 * rng.random((2, 18, 46, 54), dtype=np.float32) ** 6).astype(np.float32))
    for via_map in ("0", "1"):
        monkeypatch.setenv("OPP_GENERIC_VIA_MAP", via_map)
        for (oh, ow, k, variant) in [(300, 400, 17, 0), (369, 433, 9, 0), (47, 55, 3, 0), (60, 70, 5, 0), (97, 100, 13, 0), (300, 400, 25, 1),
                                     (333, 431, 31, 0), (200, 433, 41, 0), (368, 433, 17, 0), (150, 150, 7, 1)]:
            eng = Engine(46, 54, oh, ow, gauss_kernel_size=k, max_batch=2, max_peaks_per_part=1024, max_cands_per_limb=4096, max_humans=512, variant=variant)
            assert eng.peak_kernel() == "generic"
            H.run_and_check(eng, Oracle(46, 54, oh, ow, k, variant=variant), conf, paf, "generic %dx%d k=%d variant %d via_map=%s" % (oh, ow, k, variant, via_map))
            eng.close()


def test_std_sort_emulation_on_the_device(mods):
    """The limb kernel's two forms of std::sort(greater on score) - the parallel one (partition rounds by warps + stable rank
    sort) and the sequential emulation - run directly on candidate lists through opp_debug_sort, against the oracle's
    restatement of libstdc++'s introsort (pinned to the real std::sort on the CPU): heavy ties, few distinct values, sorted /
    reversed / organ-pipe inputs, and Musser's median-of-3 adversary, which drives ranges into the depth limit (heap sort)."""
    import ctypes as C
    Engine, Oracle, H = mods
    from openpose_plus_b200 import _capi as capi
    from oracle.oracle import CAND_DT
    eng = Engine(46, 54, max_batch=1)

    def killer(n):
        k, a = n // 2, [0] * n
        for i in range(1, k + 1):
            if i % 2:
                a[i - 1], a[i] = i, k + i
            a[k + i - 1] = 2 * i
        return np.array(a, np.float32)

    def check(scores, threads):
        n = len(scores)
        want = np.zeros(n, CAND_DT)
        want["idx1"], want["score"] = np.arange(n), scores
        want = Oracle.std_sort_desc(want)["idx1"]
        for mode in (0, 1):
            c = np.zeros(n, capi.CONN_DT)
            c["cid1"], c["score"] = np.arange(n), scores
            eng._check(eng.L.opp_debug_sort(eng.h, c.ctypes.data_as(C.c_void_p), n, mode, threads))
            assert np.array_equal(c["cid1"], want), (n, mode, threads)

    rng = np.random.default_rng(17)
    for n in (0, 1, 2, 16, 17, 18, 33, 64, 100, 257, 1000, 1024, 4096):
        for nd in (max(1, n), max(1, n // 3), 5, 2, 1):
            check((rng.integers(0, nd, n) / 7).astype(np.float32), 192)
    for n in (300, 1024, 2048):
        for arr in (np.arange(n), np.arange(n)[::-1], np.concatenate([np.arange(n // 2), np.arange(n // 2)[::-1]]), np.repeat(np.arange(n // 8), 8),
                    killer(n), -killer(n), np.floor(killer(n) / 3)):
            check(np.asarray(arr, np.float32), 256 if n == 1024 else 192)
    check((rng.integers(0, 40, 777) / 7).astype(np.float32), 32)   # a single warp takes every range in turn
    eng.close()


def test_dense_noise_frame(mods):
    """SURVEY 8(d)'s capacity-sizing frame: uniform-random maps (7 847 peaks, ~195 k candidate pairs and ~1 700 accepted
    candidates per limb, four limbs with tied scores, 1 300 partial humans, 99 merges).  Far beyond the staging areas of the
    assembly (connections and peaks are then read limb by limb / from global memory) and, with the second capacity set,
    beyond the shared-memory candidate buffers (ordered compaction, sequential std::sort emulation): every stage bit-exact."""
    Engine, Oracle, H = mods
    conf, paf = synth.noise_frame(3)
    orc = Oracle(46, 54, 368, 432, 17)
    for (capP, capC, capH) in ((512, 2048, 1536), (512, 8192, 1536)):
        eng = Engine(46, 54, max_batch=1, max_peaks_per_part=capP, max_cands_per_limb=capC, max_humans=capH)
        humans, counts, flags = H.run_and_check(eng, orc, conf[None], paf[None], "dense noise caps %d/%d/%d" % (capP, capC, capH))
        assert counts[0] > 500
        eng.close()


def test_values_around_the_peak_threshold(mods):
    """The peak kernel skips blocks that provably stay below THRESH_HEAT; maps whose smoothed maxima sit
    just below / just above 0.05, flat backgrounds near it, and isolated spikes must still be bit-exact."""
    Engine, Oracle, H = mods
    eng, orc = Engine(46, 54, max_batch=8, max_peaks_per_part=512, max_cands_per_limb=8192, max_humans=512), Oracle(46, 54, 368, 432, 17)
    conf, paf = synth.render_batch(8, n_people=6, seed0=600)
    rng = np.random.default_rng(9)
    for f, target in enumerate([0.0499, 0.05, 0.0501, 0.0505, 0.051, 0.052, 0.055, 0.06]):
        # scale so that the MEDIAN blob's smoothed maximum lands on `target`: about half the blobs pass the threshold
        sm = orc.run(conf[f], paf[f], maps=True)["smoothed"][:18]
        peaks = np.sort(sm.reshape(18, -1).max(axis=1))
        conf[f, :18] *= np.float32(target / peaks[9])
    H.run_and_check(eng, orc, conf, paf, "scaled")
    conf2, paf2 = synth.render_batch(8, n_people=4, seed0=700)
    # flat backgrounds just BELOW the threshold (a flat background above it makes every pixel a peak) + isolated spikes
    for f, bg in enumerate([0.0499, 0.04999, 0.049999, 0.0495, 0.049, 0.045, 0.03, 0.0]):
        conf2[f, :18] = np.maximum(conf2[f, :18], np.float32(bg))
        ys, xs = rng.integers(0, 46, 12), rng.integers(0, 54, 12)
        conf2[f, rng.integers(0, 18, 12), ys, xs] = rng.uniform(0.04, 0.9, 12).astype(np.float32)
    H.run_and_check(eng, orc, conf2, paf2, "background")


def test_fast_path_all_kernel_sizes_and_scale_4(mods):
    """Every (scale, kernel size) the integer-scale peak kernel is instantiated for: scale 8 with k = 1..17
    and scale 4 with k = 1..9 (k <= 5 use OpenCV's symmetric-small row form, k = 1 is a copy).  Small kernels
    on replicated maps tie whole plateaus, hence few people and large capacities."""
    Engine, Oracle, H = mods
    for (scale, k, people) in [(8, 15, 5), (8, 11, 5), (8, 7, 4), (8, 5, 1), (4, 9, 5), (4, 7, 5), (4, 5, 4), (4, 3, 1), (4, 1, 1)]:
        conf, paf = synth.render_batch(2, n_people=people, seed0=800 + k)
        oh, ow = 46 * scale, 54 * scale
        eng = Engine(46, 54, oh, ow, gauss_kernel_size=k, max_batch=2, max_peaks_per_part=1024, max_cands_per_limb=65536, max_humans=2048)
        orc = Oracle(46, 54, oh, ow, k)
        H.run_and_check(eng, orc, conf, paf, "fast x%d k=%d" % (scale, k))
        eng.close()


def test_fast_path_wide_kernels(mods):
    """S < R <= 2S: five neighbour cells instead of three (k = 19..33 at x8 - the Python graph's 25 among them - and
    k = 11..17 at x4), on rendered frames and on tiny noisy maps where every pixel is within reach of a border (the
    REFLECT_101 operands at distance exactly S and 2S), single- and multi-tile."""
    Engine, Oracle, H = mods
    conf, paf = synth.render_batch(3, n_people=6, seed0=900)
    for (scale, k) in [(8, 19), (8, 21), (8, 23), (8, 25), (8, 27), (8, 29), (8, 31), (8, 33), (4, 11), (4, 13), (4, 15), (4, 17)]:
        oh, ow = 46 * scale, 54 * scale
        eng, orc = Engine(46, 54, oh, ow, gauss_kernel_size=k, max_batch=3, max_peaks_per_part=512), Oracle(46, 54, oh, ow, k)
        assert eng.peak_kernel() == "fast", (scale, k)
        H.run_and_check(eng, orc, conf, paf, "wide x%d k=%d" % (scale, k))
        eng.close()
    rng = np.random.default_rng(12)
    for (fh, fw, scale, k) in [(3, 3, 8, 33), (3, 3, 8, 25), (3, 7, 8, 25), (5, 4, 8, 19), (4, 3, 8, 31), (9, 31, 8, 25), (33, 4, 8, 33),
                               (3, 3, 4, 17), (6, 5, 4, 11), (13, 17, 4, 13), (60, 70, 8, 25), (20, 120, 8, 27), (64, 61, 8, 25)]:
        n = 3
        conf = (rng.random((n, 19, fh, fw), dtype=np.float32) ** (3 if fh * fw < 1000 else 12)).astype(np.float32)  # large maps: sparser, within the capacities
        paf = (rng.random((n, 38, fh, fw), dtype=np.float32) * 2 - 1).astype(np.float32)
        conf[1] *= 0.04
        eng = Engine(fh, fw, fh * scale, fw * scale, gauss_kernel_size=k, max_batch=n, max_peaks_per_part=1024, max_cands_per_limb=65536, max_humans=2048)
        orc = Oracle(fh, fw, fh * scale, fw * scale, k)
        assert eng.peak_kernel() == "fast", (fh, fw, scale, k)
        H.run_and_check(eng, orc, conf, paf, "wide %dx%d x%d k=%d" % (fh, fw, scale, k))
        eng.close()
    # maps of two cells along an axis cannot hold the five-cell window: the replication-aware generic kernel takes them
    eng = Engine(2, 5, 16, 40, gauss_kernel_size=25, max_batch=1)
    assert eng.peak_kernel() == "generic_rep"
    eng.close()


def test_wide_kernels_fast_and_generic_agree(mods, monkeypatch):
    """k = 25 under both border rules through the fast kernel and, with OPP_FORCE_GENERIC=1, through the
    replication-aware generic kernel it replaced: both bit-exact against the oracle, hence against each other.  Noisy
    maps reaching into every border (where the two border rules differ)."""
    Engine, Oracle, H = mods
    from openpose_plus_b200 import _capi as capi
    rng = np.random.default_rng(13)
    conf, paf = synth.render_batch(3, n_people=8, seed0=910)
    conf[:, :18] = np.maximum(conf[:, :18], (0.4 * rng.random((3, 18, 46, 54), dtype=np.float32) ** 4).astype(np.float32))
    for variant in (capi.VARIANT_CPP, capi.VARIANT_PYTHON):
        orc = Oracle(46, 54, 368, 432, 25, variant=variant)
        for force in ("0", "1"):
            monkeypatch.setenv("OPP_FORCE_GENERIC", force)
            eng = Engine(46, 54, gauss_kernel_size=25, max_batch=3, max_peaks_per_part=512, max_cands_per_limb=8192, max_humans=512, variant=variant)
            assert eng.peak_kernel() == ("generic_rep" if force == "1" else "fast")
            H.run_and_check(eng, orc, conf, paf, "k=25 variant %d force_generic=%s" % (variant, force))
            if force == "0":  # and with the up-sampled maps materialised by the same kernel (the fused store path)
                import torch
                cu, pu = torch.empty((3, 19, 368, 432), device="cuda"), torch.empty((3, 38, 368, 432), device="cuda")
                H.run_and_check(eng, orc, conf, paf, "k=25 variant %d, fused store" % variant, conf_up=cu, paf_up=pu)
                o = orc.run(conf[2], paf[2], maps=True)
                assert np.array_equal(cu[2].cpu().numpy(), o["conf_up"]) and np.array_equal(pu[2].cpu().numpy(), o["paf_up"])
            eng.close()


def test_small_and_odd_geometries_on_the_fast_path(mods):
    """Tiny feature maps (down to 2x2) with noise: every pixel is within reach of an image border, so the
    REFLECT_101 special taps, the halo clamps and the single-tile / multi-tile logic are all exercised."""
    Engine, Oracle, H = mods
    rng = np.random.default_rng(11)
    for (fh, fw, scale, k) in [(2, 2, 8, 17), (2, 5, 8, 17), (3, 3, 8, 13), (5, 7, 8, 17), (7, 3, 8, 9), (9, 31, 8, 17), (33, 4, 8, 17),
                               (2, 2, 4, 9), (6, 5, 4, 7), (13, 17, 4, 9), (60, 70, 8, 17), (20, 120, 8, 17)]:
        n = 3
        conf = (rng.random((n, 19, fh, fw), dtype=np.float32) ** 3).astype(np.float32)      # sparse-ish noise, some values above threshold
        paf = (rng.random((n, 38, fh, fw), dtype=np.float32) * 2 - 1).astype(np.float32)
        conf[1] *= 0.04                                                                      # one frame entirely below the threshold
        eng = Engine(fh, fw, fh * scale, fw * scale, gauss_kernel_size=k, max_batch=n, max_peaks_per_part=1024, max_cands_per_limb=65536, max_humans=2048)
        orc = Oracle(fh, fw, fh * scale, fw * scale, k)
        H.run_and_check(eng, orc, conf, paf, "%dx%d x%d k=%d" % (fh, fw, scale, k))
        eng.close()


def test_python_path_variant(mods):
    """OPP_VARIANT_PYTHON: the semantics of the reference's Python graph (post_process.py:13-37: CDF-derived 25-tap
    kernel, zero padding) with pafprocess-style grouping (humans indexed by position), bit-exact against this repo's
    restatement of it (oracle variant 1; unpinned, see oracle/opp_oracle.h).  Crowded frames with the neck-nose limb
    dropped give 30+ merges per frame, where the two grouping variants differ."""
    Engine, Oracle, H = mods
    from openpose_plus_b200 import _capi as capi
    conf, paf = synth.render_batch(4, n_people=5, seed0=100)
    eng, orc = Engine(46, 54, gauss_kernel_size=25, max_batch=6, max_humans=256, variant=capi.VARIANT_PYTHON), Oracle(46, 54, 368, 432, 25, variant=1)
    H.run_and_check(eng, orc, conf, paf, "python variant")
    fr = [synth.render_frame(200 + i, n_people=30 + 2 * i, drop_limbs=(12,) if i % 2 else ()) for i in range(6)]
    conf, paf = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])
    humans, counts, flags = H.run_and_check(eng, orc, conf, paf, "python variant, crowded")
    assert not (flags & (capi.FLAG_UB_STALE_INDEX | capi.FLAG_UB_ERASE_PAST_END)).any()   # positions are never stale
    # and it is a different result from the C++ path's on the same maps
    cpp = Engine(46, 54, gauss_kernel_size=25, max_batch=6, max_humans=256)
    _, counts_cpp, _ = cpp.process(conf, paf)
    assert not np.array_equal(counts, counts_cpp)
    # other kernel sizes / geometries, maps touching the border (zero padding matters there)
    conf, paf = synth.render_batch(2, n_people=8, seed0=410)
    conf[:, :18, :3, :] = np.maximum(conf[:, :18, :3, :], 0.3 * np.random.default_rng(5).random((2, 18, 3, 54), dtype=np.float32))
    for (oh, ow, k) in [(368, 432, 17), (300, 400, 25), (46 * 4, 54 * 4, 9), (92, 108, 3)]:
        eng, orc = Engine(46, 54, oh, ow, gauss_kernel_size=k, max_batch=2, max_peaks_per_part=512, variant=capi.VARIANT_PYTHON), Oracle(46, 54, oh, ow, k, variant=1)
        H.run_and_check(eng, orc, conf, paf, "python variant %dx%d k=%d" % (oh, ow, k))

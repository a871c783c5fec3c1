"""TEST INFRASTRUCTURE -- NOT PRODUCT CODE.

ctypes view of the CPU oracle:

* ``Oracle``     -- oracle/liborc.so, the instrumented plain-C restatement (opp_oracle.c) of the
  reference path src/paf.cpp:38-57 -> src/post-process.h -> src/paf.cpp:79-311, every intermediate
  exposed.
* ``Reference``  -- oracle/_ref/libopp_ref.so, the reference's own unmodified src/paf.cpp compiled
  by oracle/Makefile against stand-in headers (final human_t lists only).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / ``--impl reference`` legs may
import this module.  Nothing under openpose_plus_b200/ does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))

N_PARTS, N_PAIRS, N_HEAT, N_PAF = 18, 19, 19, 38

PEAK_DT = np.dtype([("part_id", "<i4"), ("x", "<i4"), ("y", "<i4"), ("score", "<f4"), ("id", "<i4")])
CAND_DT = np.dtype([("idx1", "<i4"), ("idx2", "<i4"), ("score", "<f4"), ("etc", "<f4")])
CONN_DT = np.dtype([("cid1", "<i4"), ("cid2", "<i4"), ("score", "<f4")])
HREF_DT = np.dtype([("id", "<i4"), ("parts", "<i4", (N_PARTS,)), ("score", "<f4"), ("n_parts", "<i4")])
PART_DT = np.dtype([("has_value", "u1"), ("pad", "u1", (3,)), ("x", "<f4"), ("y", "<f4"), ("score", "<f4")])
HUMAN_DT = np.dtype([("parts", PART_DT, (N_PARTS,)), ("score", "<f4")])
assert HUMAN_DT.itemsize == 292 and HREF_DT.itemsize == 84 and PEAK_DT.itemsize == 20

FLAG_UB_STALE_INDEX, FLAG_UB_PEAK_INDEX, FLAG_UB_ERASE_PAST_END = 1, 2, 4


def build(force=False):
    """make -C oracle: liborc.so always; _ref/ only where /root/reference exists (make decides what is stale).
    On the GPU box there is no reference tree and usually nothing to rebuild: the prebuilt files travel."""
    have = os.path.exists(os.path.join(HERE, "liborc.so"))
    try:
        subprocess.run(["make", "-C", HERE] + (["-B"] if force else []), check=True, stdout=subprocess.DEVNULL, stderr=subprocess.PIPE)
    except (subprocess.CalledProcessError, FileNotFoundError) as e:
        if not have:
            raise RuntimeError("oracle: build failed and no prebuilt liborc.so: %s" % getattr(e, "stderr", e))


def _as(ptr, n, dt):
    if n <= 0 or not ptr:
        return np.zeros(0, dt)
    buf = (C.c_char * (n * dt.itemsize)).from_address(ptr)
    return np.frombuffer(buf, dtype=dt, count=n).copy()


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a, a.ctypes.data_as(C.POINTER(C.c_float))


class Oracle:
    _lib = None

    @classmethod
    def lib(cls):
        if cls._lib is None:
            build()
            L = C.CDLL(os.path.join(HERE, "liborc.so"))
            L.orc_create.restype = C.c_void_p
            L.orc_create.argtypes = [C.c_int] * 5
            L.orc_destroy.argtypes = [C.c_void_p]
            for f in ("orc_run", "orc_run_lazy"):
                getattr(L, f).argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float)]
                getattr(L, f).restype = C.c_int
            for f in ("orc_conf_up", "orc_paf_up", "orc_smoothed", "orc_pooled", "orc_peaks", "orc_hrefs", "orc_humans"):
                getattr(L, f).restype = C.c_void_p
                getattr(L, f).argtypes = [C.c_void_p]
            for f in ("orc_cands_unsorted", "orc_cands_sorted", "orc_conns"):
                getattr(L, f).restype = C.c_void_p
                getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            for f in ("orc_n_peaks", "orc_n_incomplete", "orc_n_merges", "orc_n_humans", "orc_flags"):
                getattr(L, f).restype = C.c_int
                getattr(L, f).argtypes = [C.c_void_p]
            for f in ("orc_n_pairs_scored", "orc_n_cands", "orc_has_score_ties", "orc_n_conns"):
                getattr(L, f).restype = C.c_int
                getattr(L, f).argtypes = [C.c_void_p, C.c_int]
            L.orc_gauss_kernel.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_float)]
            L.orc_resize_area_up.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float), C.c_int, C.c_int]
            L.orc_gauss_blur.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_double, C.POINTER(C.c_float)]
            L.orc_max_pool_3x3.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.POINTER(C.c_float)]
            L.orc_std_sort_desc.argtypes = [C.c_void_p, C.c_int]
            L.orc_resize_coeffs.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_float)]
            L.orc_resize_coeffs.restype = C.c_int
            L.orc_set_variant.argtypes = [C.c_void_p, C.c_int]
            L.orc_set_variant.restype = None
            L.orc_cdf_kernel.argtypes = [C.c_int, C.c_double, C.POINTER(C.c_float)]
            L.orc_smooth_zero_pad.argtypes = [C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float), C.POINTER(C.c_float)]
            cls._lib = L
        return cls._lib

    def __init__(self, feat_h, feat_w, out_h, out_w, ksize=17, variant=0):
        """variant 0 = the C++ path (src/paf.cpp, pinned); 1 = the Python path's semantics (unpinned, see opp_oracle.h)."""
        self.L = self.lib()
        self.h, self.w, self.H, self.W, self.ksize = feat_h, feat_w, out_h, out_w, ksize
        self.ctx = self.L.orc_create(feat_h, feat_w, out_h, out_w, ksize)
        if not self.ctx:
            raise ValueError("oracle: unsupported geometry/kernel size")
        self.L.orc_set_variant(self.ctx, variant)

    def __del__(self):
        if getattr(self, "ctx", None):
            self.L.orc_destroy(self.ctx)
            self.ctx = None

    def run(self, conf, paf, lazy=False, maps=False):
        """One frame -> dict of every intermediate (numpy copies)."""
        conf, pc = _f32(conf)
        paf, pp = _f32(paf)
        assert conf.shape == (N_HEAT, self.h, self.w) and paf.shape == (N_PAF, self.h, self.w)
        L, ctx = self.L, self.ctx
        n = (L.orc_run_lazy if lazy else L.orc_run)(ctx, pc, pp)
        out = {
            "n_humans": n,
            "peaks": _as(L.orc_peaks(ctx), L.orc_n_peaks(ctx), PEAK_DT),
            "cands_unsorted": [_as(L.orc_cands_unsorted(ctx, p), L.orc_n_cands(ctx, p), CAND_DT) for p in range(N_PAIRS)],
            "cands_sorted": [_as(L.orc_cands_sorted(ctx, p), L.orc_n_cands(ctx, p), CAND_DT) for p in range(N_PAIRS)],
            "ties": [L.orc_has_score_ties(ctx, p) for p in range(N_PAIRS)],
            "n_pairs": [L.orc_n_pairs_scored(ctx, p) for p in range(N_PAIRS)],
            "conns": [_as(L.orc_conns(ctx, p), L.orc_n_conns(ctx, p), CONN_DT) for p in range(N_PAIRS)],
            "hrefs": _as(L.orc_hrefs(ctx), n, HREF_DT),
            "humans": _as(L.orc_humans(ctx), n, HUMAN_DT),
            "n_incomplete": L.orc_n_incomplete(ctx),
            "n_merges": L.orc_n_merges(ctx),
            "flags": L.orc_flags(ctx),
        }
        if maps:
            px = self.H * self.W
            out["conf_up"] = _as(L.orc_conf_up(ctx), N_HEAT * px, np.dtype("<f4")).reshape(N_HEAT, self.H, self.W)
            out["smoothed"] = _as(L.orc_smoothed(ctx), N_HEAT * px, np.dtype("<f4")).reshape(N_HEAT, self.H, self.W)
            out["pooled"] = _as(L.orc_pooled(ctx), N_HEAT * px, np.dtype("<f4")).reshape(N_HEAT, self.H, self.W)
            if not lazy:
                out["paf_up"] = _as(L.orc_paf_up(ctx), N_PAF * px, np.dtype("<f4")).reshape(N_PAF, self.H, self.W)
        return out

    # stand-alone stages ---------------------------------------------------------------------
    @classmethod
    def gauss_kernel(cls, k, sigma=3.0):
        t = np.zeros(k, np.float32)
        if cls.lib().orc_gauss_kernel(k, sigma, t.ctypes.data_as(C.POINTER(C.c_float))):
            raise ValueError("bad kernel size")
        return t

    @classmethod
    def cdf_kernel(cls, k, nsig=3.0):
        t = np.zeros(k, np.float32)
        if cls.lib().orc_cdf_kernel(k, nsig, t.ctypes.data_as(C.POINTER(C.c_float))):
            raise ValueError("bad kernel size")
        return t

    @classmethod
    def smooth_zero_pad(cls, plane, taps):
        plane, p = _f32(plane)
        taps, pt = _f32(taps)
        out = np.empty_like(plane)
        if cls.lib().orc_smooth_zero_pad(p, plane.shape[0], plane.shape[1], len(taps), pt, out.ctypes.data_as(C.POINTER(C.c_float))):
            raise ValueError("bad kernel size")
        return out

    @classmethod
    def resize_area(cls, plane, H, W):
        plane, p = _f32(plane)
        out = np.empty((H, W), np.float32)
        if cls.lib().orc_resize_area_up(p, plane.shape[0], plane.shape[1], out.ctypes.data_as(C.POINTER(C.c_float)), H, W):
            raise ValueError("down-sampling is not restated")
        return out

    @classmethod
    def resize_coeffs(cls, ssize, dsize):
        ofs = np.zeros(dsize, np.int32)
        alpha = np.zeros((dsize, 2), np.float32)
        dmax = cls.lib().orc_resize_coeffs(ssize, dsize, ofs.ctypes.data_as(C.POINTER(C.c_int)), alpha.ctypes.data_as(C.POINTER(C.c_float)))
        return ofs, alpha, dmax

    @classmethod
    def gauss_blur(cls, plane, k, sigma=3.0):
        plane, p = _f32(plane)
        out = np.empty_like(plane)
        if cls.lib().orc_gauss_blur(p, plane.shape[0], plane.shape[1], k, sigma, out.ctypes.data_as(C.POINTER(C.c_float))):
            raise ValueError("bad kernel size")
        return out

    @classmethod
    def max_pool(cls, plane):
        plane, p = _f32(plane)
        out = np.empty_like(plane)
        cls.lib().orc_max_pool_3x3(p, plane.shape[0], plane.shape[1], out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    @classmethod
    def std_sort_desc(cls, cands):
        v = np.ascontiguousarray(cands, dtype=CAND_DT).copy()
        cls.lib().orc_std_sort_desc(v.ctypes.data, len(v))
        return v


def ref_available(fast=False):
    return os.path.exists(os.path.join(HERE, "_ref", "libopp_ref_fast.so" if fast else "libopp_ref.so"))


class Reference:
    """The reference's own paf_processor (create_paf_processor, include/openpose-plus.hpp:54-64)."""

    _libs = {}

    @classmethod
    def lib(cls, fast=False):
        if fast not in cls._libs:
            build()
            L = C.CDLL(os.path.join(HERE, "_ref", "libopp_ref_fast.so" if fast else "libopp_ref.so"))
            L.ref_create.restype = C.c_void_p
            L.ref_create.argtypes = [C.c_int] * 5
            L.ref_destroy.argtypes = [C.c_void_p]
            L.ref_run.restype = C.c_int
            L.ref_run.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_void_p, C.c_int]
            L.ref_time_frames.restype = C.c_double
            L.ref_time_frames.argtypes = [C.c_int] * 5 + [C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_long)]
            L.ref_std_sort_desc.argtypes = [C.c_void_p, C.c_int]
            L.ref_latency_frames.restype = None
            L.ref_latency_frames.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_double)]
            L.ref_pool_create.restype = C.c_void_p
            L.ref_pool_create.argtypes = [C.c_int] * 6
            L.ref_pool_destroy.argtypes = [C.c_void_p]
            L.ref_pool_run.restype = C.c_double
            L.ref_pool_run.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.POINTER(C.c_float), C.c_int, C.POINTER(C.c_long)]
            cls._libs[fast] = L
        return cls._libs[fast]

    def __init__(self, feat_h, feat_w, out_h, out_w, ksize=17, fast=False):
        self.L = self.lib(fast)
        self.geom = (feat_h, feat_w, out_h, out_w, ksize)
        self.p = self.L.ref_create(feat_h, feat_w, out_h, out_w, ksize)

    def __del__(self):
        if getattr(self, "p", None):
            self.L.ref_destroy(self.p)
            self.p = None

    def run(self, conf, paf, cap=4096):
        conf, pc = _f32(conf)
        paf, pp = _f32(paf)
        out = np.zeros(cap, HUMAN_DT)
        n = self.L.ref_run(self.p, pc, pp, out.ctypes.data, cap)
        assert n <= cap
        return out[:n].copy()

    def latency_ms(self, conf, paf, iters):
        """Per-call wall time (ms) of `iters` one-frame calls on this processor, one thread, stdout silenced."""
        conf, pc = _f32(conf)
        paf, pp = _f32(paf)
        out = np.zeros(iters, np.float64)
        self.L.ref_latency_frames(self.p, pc, pp, conf.shape[0], self.geom[0], self.geom[1], iters, out.ctypes.data_as(C.POINTER(C.c_double)))
        return out

    @classmethod
    def time_frames(cls, geom, conf, paf, repeat=1, threads=1, fast=True):
        """Wall seconds for len(conf)*repeat frames on `threads` independent processors."""
        L = cls.lib(fast)
        conf, pc = _f32(conf)
        paf, pp = _f32(paf)
        tot = C.c_long(0)
        s = L.ref_time_frames(*geom, pc, pp, conf.shape[0], repeat, threads, C.byref(tot))
        return s, tot.value

    @classmethod
    def std_sort_desc(cls, cands):
        v = np.ascontiguousarray(cands, dtype=CAND_DT).copy()
        cls.lib().ref_std_sort_desc(v.ctypes.data, len(v))
        return v


class ReferencePool:
    """`threads` reference processors kept alive across calls; run() returns wall seconds for the frames given."""

    def __init__(self, geom, threads, fast=True):
        self.L = Reference.lib(fast)
        self.p = self.L.ref_pool_create(*geom, threads)
        self.threads = threads

    def run(self, conf, paf):
        conf, pc = _f32(conf)
        paf, pp = _f32(paf)
        tot = C.c_long(0)
        return self.L.ref_pool_run(self.p, pc, pp, conf.shape[0], C.byref(tot)), tot.value

    def __del__(self):
        if getattr(self, "p", None):
            self.L.ref_pool_destroy(self.p)
            self.p = None

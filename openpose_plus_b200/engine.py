"""Host-side engine over the C-ABI: one handle per GPU, batches in, skeletons out.

This is the batched/streaming driver around the reference's one-frame-per-call contract
(/root/reference examples/pose_detector.cpp:96-102): `process` is the synchronous call,
`submit`/`wait` keep several batches in flight (copy / compute overlap on separate streams).
"""
import ctypes as C

import numpy as np

from . import _capi as capi


def _ptr(a):
    """Raw address of a numpy array, a torch tensor or anything exposing data_ptr()."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        try:  # ~2.5x cheaper than a.ctypes.data; matters on the one-frame latency path (five arrays per call)
            return C.addressof(C.c_char.from_buffer(a))
        except (TypeError, ValueError, BufferError):  # read-only or exotic buffers
            return a.ctypes.data
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    raise TypeError("expected numpy array or tensor, got %r" % type(a))


def _is_device(a):
    return (not isinstance(a, np.ndarray)) and hasattr(a, "is_cuda") and bool(a.is_cuda)


class Engine:
    def __init__(self, feat_h, feat_w, out_h=None, out_w=None, gauss_kernel_size=17, max_batch=64, device=-1,
                 max_peaks_per_part=128, max_cands_per_limb=1024, max_humans=128, n_slots=3, variant=capi.VARIANT_CPP):
        self.L = capi.lib()
        out_h = 8 * feat_h if out_h is None else out_h
        out_w = 8 * feat_w if out_w is None else out_w
        cfg = capi.Config()
        self.L.opp_config_default(C.byref(cfg), feat_h, feat_w, out_h, out_w, gauss_kernel_size)
        cfg.max_batch, cfg.device = max_batch, device
        cfg.max_peaks_per_part, cfg.max_cands_per_limb, cfg.max_humans, cfg.n_slots = max_peaks_per_part, max_cands_per_limb, max_humans, n_slots
        cfg.variant = variant
        self.cfg = cfg
        self.h = C.c_void_p()
        rc = self.L.opp_create(C.byref(cfg), C.byref(self.h))
        if rc != capi.OK:
            raise capi.OppError(rc, (self.L.opp_last_error(None) or b"").decode())
        self.device = int(self.L.opp_device(self.h))
        self.feat = (feat_h, feat_w)
        self.out = (out_h, out_w)
        self.max_batch, self.max_humans = max_batch, max_humans
        self._pending = {}

    def close(self):
        if getattr(self, "h", None):
            self.L.opp_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != capi.OK:
            raise capi.OppError(rc, (self.L.opp_last_error(self.h) or b"").decode())

    # ------------------------------------------------------------------------------------------
    def submit(self, conf, paf, layout=capi.LAYOUT_CHW, conf_up=None, paf_up=None, up_layout=capi.LAYOUT_CHW, out=None):
        """conf [n,19,h,w] / paf [n,38,h,w] (or channels-last), numpy (host) or CUDA tensors.
        Returns a ticket for wait().  `out` may carry pre-allocated (humans, n_humans, flags): numpy arrays
        (host results) or CUDA uint8/int32 tensors of the same byte sizes (results stay on the device)."""
        dev = _is_device(conf)
        if not dev:
            conf = np.ascontiguousarray(conf, np.float32)
            paf = np.ascontiguousarray(paf, np.float32)
        n = int(conf.shape[0])
        if out is None:
            out = (np.zeros((n, self.max_humans), capi.HUMAN_DT), np.zeros(n, np.int32), np.zeros(n, np.int32))
        humans, counts, flags = out
        b = capi.Batch()
        b.conf, b.paf, b.n_frames = _ptr(conf), _ptr(paf), n
        b.in_mem = capi.MEM_DEVICE if dev else capi.MEM_HOST
        b.in_layout, b.out_mem = layout, (capi.MEM_DEVICE if _is_device(humans) else capi.MEM_HOST)
        b.humans, b.n_humans, b.frame_flags = _ptr(humans), _ptr(counts), _ptr(flags)
        b.conf_up, b.paf_up, b.up_layout = _ptr(conf_up), _ptr(paf_up), up_layout
        t = C.c_int(-1)
        self._check(self.L.opp_submit(self.h, C.byref(b), C.byref(t)))
        self._pending[t.value] = (conf, paf, conf_up, paf_up, humans, counts, flags)
        return t.value

    def wait(self, ticket):
        self._check(self.L.opp_wait(self.h, ticket))
        _, _, _, _, humans, counts, flags = self._pending.pop(ticket)
        return humans, counts, flags

    def process(self, conf, paf, **kw):
        """Synchronous: returns (humans [n,max_humans] HUMAN_DT, n_humans [n], flags [n])."""
        return self.wait(self.submit(conf, paf, **kw))

    def last_batch_ms(self, ticket):
        return float(self.L.opp_last_batch_ms(self.h, ticket))

    def launch_count(self):
        return int(self.L.opp_launch_count(self.h))

    # ---- intermediates of the last batch on a ticket (parity tests) ------------------------------
    def debug_peaks(self, ticket, frame, cap=None):
        cap = cap or 18 * self.cfg.max_peaks_per_part
        buf = np.zeros(cap, capi.PEAK_DT)
        n = self.L.opp_debug_fetch(self.h, ticket, capi.DBG_PEAKS, frame, 0, buf.ctypes.data, cap)
        if n < 0:
            raise capi.OppError(-1, "debug_fetch(peaks) failed")
        return buf[:min(n, cap)].copy()

    def debug_conns(self, ticket, frame, pair_id):
        cap = self.cfg.max_peaks_per_part
        buf = np.zeros(cap, capi.CONN_DT)
        n = self.L.opp_debug_fetch(self.h, ticket, capi.DBG_CONNS, frame, pair_id, buf.ctypes.data, cap)
        if n < 0:
            raise capi.OppError(-1, "debug_fetch(conns) failed")
        return buf[:min(n, cap)].copy()

    def debug_parts(self, ticket, frame, human):
        buf = np.zeros(18, np.int32)
        if self.L.opp_debug_fetch(self.h, ticket, capi.DBG_PARTS, frame, human, buf.ctypes.data, 18) < 0:
            raise capi.OppError(-1, "debug_fetch(parts) failed")
        return buf

    def debug_counts(self, ticket, frame):
        buf = np.zeros(4, np.int32)
        if self.L.opp_debug_fetch(self.h, ticket, capi.DBG_COUNTS, frame, 0, buf.ctypes.data, 4) < 0:
            raise capi.OppError(-1, "debug_fetch(counts) failed")
        return buf

    def resize(self, src, dst, layout=capi.LAYOUT_CHW, stream=None):
        """Stand-alone up-sampling of device maps src [n,C,h,w] -> dst ([n,C,H,W] or [n,H,W,C])."""
        n, ch = int(src.shape[0]), int(src.shape[1])
        self._check(self.L.opp_resize_device(self.h, _ptr(src), ch, n, _ptr(dst), layout, stream))

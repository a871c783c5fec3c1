"""Host-side engine over the C-ABI: one handle per GPU, batches in, skeletons out.

This is the batched/streaming driver around the reference's one-frame-per-call contract
(/root/reference examples/pose_detector.cpp:96-102): `process` is the synchronous call,
`submit`/`wait` keep several batches in flight (copy / compute overlap on separate streams).
"""
import ctypes as C

import numpy as np

from . import _capi as capi
from . import dlpack


def _ptr(a):
    """Raw address of a numpy array or of any DLPack-capable / data_ptr() buffer (no validation: internal use)."""
    if a is None:
        return None
    return dlpack.resolve(a, "buffer").ptr


def _stream_handle(s):
    """cudaStream_t / cudaEvent_t as an integer: ints pass through, torch.cuda.Stream has .cuda_stream, torch.cuda.Event
    .cuda_event, cupy streams / events .ptr."""
    if s is None:
        return None
    if isinstance(s, int):
        return s
    for attr in ("cuda_stream", "cuda_event", "ptr", "handle"):
        v = getattr(s, attr, None)
        if v is not None:
            return int(v)
    raise TypeError("expected a CUDA stream / event handle, got %r" % type(s))


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
    def _bad(self, msg):
        raise capi.OppError(capi.ERR_INVALID, msg)

    def _map_in(self, a, ch, what, layout):
        """One validated input tensor -> (keepalive, address, n_frames, on_device).  numpy arrays are normalised to
        contiguous float32 (this is the one-frame latency path: a handful of cheap checks, no helper objects)."""
        h, w = self.feat
        want = (ch, h, w) if layout == capi.LAYOUT_CHW else (h, w, ch)
        if type(a) is np.ndarray:
            a = np.ascontiguousarray(a, np.float32)
            shape, dev = a.shape, False
            try:
                ptr = C.addressof(C.c_char.from_buffer(a))
            except (TypeError, ValueError, BufferError):  # read-only arrays are fine as inputs
                ptr = a.ctypes.data
        else:
            try:
                a = dlpack.resolve(a, what, dlpack.F32)
            except TypeError as e:
                self._bad(str(e))
            shape, dev, ptr = a.shape, a.on_device, a.ptr
            if dev and a.device_id not in (-1, self.device):
                self._bad("%s lives on cuda:%d, this engine on cuda:%d" % (what, a.device_id, self.device))
        if len(shape) != 4 or tuple(shape[1:]) != want:
            self._bad("%s: shape %s, expected [n, %d, %d, %d]" % (what, tuple(shape), *want))
        return a, ptr, int(shape[0]), dev

    def _buf_out(self, a, what, min_bytes, itemsize=0, device_only=False):
        """One validated output buffer -> (keepalive, address, on_device)."""
        if a is None:
            return None, None, False
        if type(a) is np.ndarray:
            if a.nbytes < min_bytes or (itemsize and a.dtype.itemsize != itemsize) or device_only:
                self._bad("%s: numpy array of %d bytes (%s); need %s%d bytes" % (what, a.nbytes, a.dtype, "device memory, " if device_only else "", min_bytes))
            try:
                return a, C.addressof(C.c_char.from_buffer(a)), False
            except (TypeError, ValueError, BufferError):
                self._bad("%s must be a writable, C-contiguous array" % what)
        try:
            b = dlpack.resolve(a, what, dlpack.I32 if itemsize == 4 else (dlpack.F32 if device_only else dlpack.ANY), writable=True)
        except TypeError as e:
            self._bad(str(e))
        if b.nbytes and b.nbytes < min_bytes:
            self._bad("%s: %d bytes, at least %d needed" % (what, b.nbytes, min_bytes))
        if device_only and not b.on_device:
            self._bad("%s must be device memory" % what)
        if b.on_device and b.device_id not in (-1, self.device):
            self._bad("%s lives on cuda:%d, this engine on cuda:%d" % (what, b.device_id, self.device))
        return b, b.ptr, b.on_device

    def submit(self, conf, paf, layout=capi.LAYOUT_CHW, conf_up=None, paf_up=None, up_layout=capi.LAYOUT_CHW, out=None,
               in_stream=None, in_event=None):
        """conf [n,19,h,w] / paf [n,38,h,w] (or channels-last): numpy arrays (host) or any DLPack-capable buffer
        (torch / cupy / jax arrays, a 'dltensor' capsule; host or CUDA memory), float32, compact.  Returns a ticket for
        wait().  `out` may carry pre-allocated (humans, n_humans, flags): numpy arrays (host results) or CUDA
        uint8/int32 buffers of the same byte sizes (results stay on the device).  Device-resident inputs written by a
        producer on its own stream: pass `in_stream` (the producer's stream) or `in_event` (recorded after its last
        write) and the hand-off is ordered on the device, without a host synchronisation."""
        conf, pc, n, dev = self._map_in(conf, capi.N_HEAT, "conf", layout)
        paf, pp, n2, dev2 = self._map_in(paf, capi.N_PAF, "paf", layout)
        if n != n2 or dev != dev2:
            self._bad("conf and paf must hold the same number of frames in the same memory kind")
        if out is None:
            out = (np.zeros((n, self.max_humans), capi.HUMAN_DT), np.zeros(n, np.int32), np.zeros(n, np.int32))
        humans, counts, flags = out
        H, W = self.out
        kh, ph, dh = self._buf_out(humans, "humans", n * self.max_humans * 292)
        kc, pcn, dc = self._buf_out(counts, "n_humans", 4 * n, 4)
        kf, pf, df = self._buf_out(flags, "frame_flags", 4 * n, 4)
        if ph is None or pcn is None or dh != dc or (pf is not None and df != dh):
            self._bad("humans, n_humans (and frame_flags) are required and must live in the same memory kind")
        kcu, pcu, _ = self._buf_out(conf_up, "conf_up", 4 * n * capi.N_HEAT * H * W, device_only=True)
        kpu, ppu, _ = self._buf_out(paf_up, "paf_up", 4 * n * capi.N_PAF * H * W, device_only=True)
        b = capi.Batch()
        b.conf, b.paf, b.n_frames = pc, pp, n
        b.in_mem = capi.MEM_DEVICE if dev else capi.MEM_HOST
        b.in_layout, b.out_mem = layout, (capi.MEM_DEVICE if dh else capi.MEM_HOST)
        b.humans, b.n_humans, b.frame_flags = ph, pcn, pf
        b.conf_up, b.paf_up, b.up_layout = pcu, ppu, up_layout
        if in_event is not None:
            b.in_sync, b.in_sync_obj = capi.SYNC_EVENT, _stream_handle(in_event)
        elif in_stream is not None:
            b.in_sync, b.in_sync_obj = capi.SYNC_STREAM, _stream_handle(in_stream)
        t = C.c_int(-1)
        self._check(self.L.opp_submit(self.h, C.byref(b), C.byref(t)))
        self._pending[t.value] = (humans, counts, flags, (conf, paf, kh, kc, kf, kcu, kpu))
        return t.value

    def wait(self, ticket):
        self._check(self.L.opp_wait(self.h, ticket))
        humans, counts, flags, held = self._pending.pop(ticket)
        for b in held:  # consumed DLPack tensors go back to their producers
            if isinstance(b, dlpack.Buffer):
                b.release()
        return humans, counts, flags

    def stream_wait(self, ticket, stream):
        """Device-side wait: `stream` will not run past this point before the batch behind `ticket` is complete."""
        self._check(self.L.opp_stream_wait_ticket(self.h, ticket, _stream_handle(stream)))

    def process(self, conf, paf, **kw):
        """Synchronous: returns (humans [n,max_humans] HUMAN_DT, n_humans [n], flags [n])."""
        return self.wait(self.submit(conf, paf, **kw))

    def last_batch_ms(self, ticket):
        return float(self.L.opp_last_batch_ms(self.h, ticket))

    def bounds_report(self):
        """(line, index, size, violations) of the first out-of-bounds access a bounds-checked build recorded since the
        last call (all 0: none); OppError on a release build.  See opp_debug_bounds_report."""
        out = (C.c_int32 * 4)()
        self._check(self.L.opp_debug_bounds_report(self.h, out))
        return tuple(out)

    def peak_kernel(self):
        """'fast' | 'generic_rep' | 'generic': the peak kernel opp_create selected (opp_peak_kernel)."""
        return self.L.opp_peak_kernel(self.h).decode()

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

"""ctypes binding of the C-ABI in include/opp_b200.h (the same entry points the C++ `paf_processor`
wrapper calls).  Loading fails loudly when the CUDA library has not been built: there is no CPU path."""
import ctypes as C
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# OPP_B200_LIB selects another build of the same library (e.g. libopp_b200_dbg.so, the bounds-checked one)
LIB_PATH = os.environ.get("OPP_B200_LIB") or os.path.join(HERE, "libopp_b200.so")

N_PARTS, N_PAIRS, N_HEAT, N_PAF = 18, 19, 19, 38
MEM_HOST, MEM_DEVICE = 0, 1
LAYOUT_CHW, LAYOUT_HWC = 0, 1
OK, ERR_INVALID, ERR_CUDA, ERR_NO_DEVICE, ERR_BUSY = 0, 1, 2, 3, 4
FLAG_PEAK_OVERFLOW, FLAG_CAND_OVERFLOW, FLAG_HUMAN_OVERFLOW = 1, 2, 4
FLAG_UB_STALE_INDEX, FLAG_UB_PEAK_INDEX, FLAG_UB_ERASE_PAST_END = 8, 16, 32
FLAG_OVERFLOW_MASK = 7
VARIANT_CPP, VARIANT_PYTHON = 0, 1
DBG_PEAKS, DBG_CONNS, DBG_PARTS, DBG_COUNTS = 0, 1, 2, 3
SYNC_NONE, SYNC_STREAM, SYNC_EVENT = 0, 1, 2
HOST_DEFAULT, HOST_WRITE_COMBINED = 0, 1

PART_DT = np.dtype([("has_value", "u1"), ("pad", "u1", (3,)), ("x", "<f4"), ("y", "<f4"), ("score", "<f4")])
HUMAN_DT = np.dtype([("parts", PART_DT, (N_PARTS,)), ("score", "<f4")])
PEAK_DT = np.dtype([("part_id", "<i4"), ("x", "<i4"), ("y", "<i4"), ("score", "<f4"), ("id", "<i4")])
CONN_DT = np.dtype([("cid1", "<i4"), ("cid2", "<i4"), ("score", "<f4")])
assert HUMAN_DT.itemsize == 292 and PEAK_DT.itemsize == 20 and CONN_DT.itemsize == 12


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in (
        "feat_h", "feat_w", "out_h", "out_w", "n_joins", "n_connections", "gauss_kernel_size", "max_batch",
        "device", "max_peaks_per_part", "max_cands_per_limb", "max_humans", "n_slots", "variant")] + [("reserved", C.c_int32 * 2)]


class Batch(C.Structure):
    _fields_ = [("conf", C.c_void_p), ("paf", C.c_void_p), ("n_frames", C.c_int32), ("in_mem", C.c_int32),
                ("in_layout", C.c_int32), ("out_mem", C.c_int32), ("humans", C.c_void_p), ("n_humans", C.c_void_p),
                ("frame_flags", C.c_void_p), ("conf_up", C.c_void_p), ("paf_up", C.c_void_p), ("up_layout", C.c_int32),
                ("in_sync", C.c_int32), ("in_sync_obj", C.c_void_p)]


assert C.sizeof(Batch) == 88 and C.sizeof(Config) == 64  # static_assert'ed on the C side too


class OppError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("opp error %d: %s" % (code, msg))
        self.code = code


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                "openpose_plus_b200: %s is missing. Build it with `python -m openpose_plus_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback for this path." % LIB_PATH)
        L = C.CDLL(LIB_PATH)
        L.opp_config_default.argtypes = [C.POINTER(Config)] + [C.c_int] * 5
        L.opp_config_default.restype = None
        L.opp_create.argtypes = [C.POINTER(Config), C.POINTER(C.c_void_p)]
        L.opp_destroy.argtypes = [C.c_void_p]
        L.opp_destroy.restype = None
        L.opp_process.argtypes = [C.c_void_p, C.POINTER(Batch)]
        L.opp_submit.argtypes = [C.c_void_p, C.POINTER(Batch), C.POINTER(C.c_int)]
        L.opp_wait.argtypes = [C.c_void_p, C.c_int]
        L.opp_last_batch_ms.argtypes = [C.c_void_p, C.c_int]
        L.opp_last_batch_ms.restype = C.c_float
        L.opp_device.argtypes = [C.c_void_p]
        L.opp_launch_count.argtypes = [C.c_void_p]
        L.opp_launch_count.restype = C.c_int64
        L.opp_peak_kernel.argtypes = [C.c_void_p]
        L.opp_peak_kernel.restype = C.c_char_p
        L.opp_host_alloc.argtypes = [C.c_size_t]
        L.opp_host_alloc.restype = C.c_void_p
        L.opp_host_free.argtypes = [C.c_void_p]
        L.opp_host_free.restype = None
        L.opp_host_alloc_ex.argtypes = [C.c_size_t, C.c_int]
        L.opp_host_alloc_ex.restype = C.c_void_p
        L.opp_host_register.argtypes = [C.c_void_p, C.c_size_t]
        L.opp_host_unregister.argtypes = [C.c_void_p]
        L.opp_stream_wait_ticket.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.opp_debug_fetch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.opp_resize_device.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.opp_resize_pair_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.opp_peaks_device.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        L.opp_timer_start.argtypes = [C.c_void_p]
        L.opp_timer_stop.argtypes = [C.c_void_p]
        L.opp_timer_stop.restype = C.c_float
        L.opp_last_error.argtypes = [C.c_void_p]
        L.opp_last_error.restype = C.c_char_p
        L.opp_version.restype = C.c_char_p
        L.opp_bench_latency.argtypes = [C.c_void_p, C.POINTER(Batch), C.c_int, C.c_void_p]
        L.opp_bench_h2d.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_float)]
        L.opp_draw_human.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_ssize_t, C.c_void_p, C.c_int]
        _lib = L
    return _lib


EXPORTS = ["opp_config_default", "opp_create", "opp_destroy", "opp_process", "opp_submit", "opp_wait",
           "opp_last_batch_ms", "opp_launch_count", "opp_peak_kernel", "opp_debug_bounds_report", "opp_debug_sort", "opp_device", "opp_host_alloc", "opp_host_free", "opp_host_alloc_ex",
           "opp_host_register", "opp_host_unregister", "opp_stream_wait_ticket", "opp_debug_fetch",
           "opp_resize_device", "opp_resize_pair_device", "opp_peaks_device", "opp_timer_start", "opp_timer_stop", "opp_last_error", "opp_version", "opp_draw_human", "opp_bench_latency", "opp_bench_h2d", "process_conf_paf"]


def pinned_empty(shape, dtype, write_combined=False):
    """numpy array over pinned host memory; freed when the last view of it is collected.  write_combined=True is for
    INPUT buffers the host only writes (see OPP_HOST_WRITE_COMBINED)."""
    import weakref
    dtype = np.dtype(dtype)
    count = int(np.prod(shape))
    n = max(count * dtype.itemsize, 1)
    L = lib()
    p = L.opp_host_alloc_ex(n, HOST_WRITE_COMBINED if write_combined else HOST_DEFAULT)
    if not p:
        raise MemoryError("opp_host_alloc(%d) failed" % n)
    buf = (C.c_char * n).from_address(p)
    weakref.finalize(buf, L.opp_host_free, p)
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)

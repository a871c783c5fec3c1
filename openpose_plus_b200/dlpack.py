"""Buffer intake for the Python host side: numpy arrays and anything that speaks DLPack.

`north_star`: "Python calls that same C-ABI with numpy or DLPack buffers".  A producer on the device side of the
reference's runner -> paf_processor seam (/root/reference src/uff-runner.cpp:208-217, include/openpose-plus.hpp:42-51) -
a torch / cupy / jax array, a TensorRT binding wrapped in a capsule - hands its buffer over through `__dlpack__`; the
capsule is consumed here with the CPython capsule API through ctypes (no torch import, no copy).  Every buffer is
validated before its address crosses the C-ABI: device, dtype, compact row-major strides and byte size.  The C-ABI takes
raw pointers and cannot check any of that itself.
"""
import ctypes as C

import numpy as np

# dlpack.h: DLDeviceType
kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
# dlpack.h: DLDataTypeCode
kDLInt, kDLUInt, kDLFloat = 0, 1, 2

F32 = ((kDLFloat, 32),)
I32 = ((kDLInt, 32), (kDLUInt, 32))
ANY = None


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]

_api = C.pythonapi
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_SetName.restype = C.c_int
_api.PyCapsule_SetName.argtypes = [C.py_object, C.c_char_p]
_USED = C.c_char_p(b"used_dltensor")  # must outlive every capsule renamed to it


class BufferError_(TypeError):
    """A buffer that must not cross the C-ABI (wrong device / dtype / strides / size)."""


class Buffer:
    """A validated buffer: raw address, where it lives, and whatever keeps it alive."""
    __slots__ = ("ptr", "on_device", "device_id", "shape", "nbytes", "_owner", "_managed")

    def __init__(self, ptr, on_device, device_id, shape, nbytes, owner, managed=None):
        self.ptr, self.on_device, self.device_id, self.shape, self.nbytes = ptr, on_device, device_id, tuple(shape), nbytes
        self._owner, self._managed = owner, managed

    def release(self):
        """Gives a consumed DLPack tensor back to its producer (DLManagedTensor.deleter)."""
        m, self._managed = self._managed, None
        if m is not None and m.contents.deleter:
            m.contents.deleter(m)
        self._owner = None

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass


def _fail(what, msg):
    raise BufferError_("%s: %s" % (what, msg))


def from_numpy(a, what, dtypes=ANY, writable=False):
    if not a.flags.c_contiguous:
        _fail(what, "numpy array must be C-contiguous")
    if writable and not a.flags.writeable:
        _fail(what, "numpy array must be writable")
    if dtypes is not None:
        code = {"f": kDLFloat, "i": kDLInt, "u": kDLUInt}.get(a.dtype.kind)
        if (code, a.dtype.itemsize * 8) not in dtypes or not a.dtype.isnative:
            _fail(what, "dtype %s not accepted" % a.dtype)
    try:  # ~2.5x cheaper than a.ctypes.data; matters on the one-frame latency path (five arrays per call)
        ptr = C.addressof(C.c_char.from_buffer(a))
    except (TypeError, ValueError, BufferError):  # read-only or exotic buffers
        ptr = a.ctypes.data
    return Buffer(ptr, False, -1, a.shape, a.nbytes, a)


def from_dlpack(obj, what, dtypes=ANY):
    """Consumes obj.__dlpack__() (or a ready 'dltensor' capsule)."""
    cap = obj
    if not (type(obj).__name__ == "PyCapsule"):
        try:
            cap = obj.__dlpack__()
        except Exception as e:  # e.g. torch tensors that require grad
            _fail(what, "__dlpack__() failed: %s" % e)
    if not _api.PyCapsule_IsValid(cap, b"dltensor"):
        _fail(what, "expected a DLPack capsule named 'dltensor' (already consumed, or a versioned capsule)")
    m = C.cast(_api.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(_DLManagedTensor))
    t = m.contents.dl_tensor
    ok = False
    try:
        if t.device.device_type in (kDLCUDA, kDLCUDAManaged):
            on_device = True
        elif t.device.device_type in (kDLCPU, kDLCUDAHost):
            on_device = False
        else:
            _fail(what, "DLPack device type %d is neither host nor CUDA memory" % t.device.device_type)
        if t.dtype.lanes != 1 or (dtypes is not None and (t.dtype.code, t.dtype.bits) not in dtypes):
            _fail(what, "DLPack dtype (code %d, %d bits, %d lanes) not accepted" % (t.dtype.code, t.dtype.bits, t.dtype.lanes))
        shape = [int(t.shape[i]) for i in range(t.ndim)]
        if t.strides:  # NULL = compact row-major
            expect = 1
            for i in range(t.ndim - 1, -1, -1):
                if shape[i] != 1 and int(t.strides[i]) != expect:
                    _fail(what, "DLPack tensor must be compact row-major (shape %s, strides %s)" % (shape, [int(t.strides[i]) for i in range(t.ndim)]))
                expect *= shape[i]
        n = 1
        for d in shape:
            n *= d
        ptr = (t.data or 0) + int(t.byte_offset)
        buf = Buffer(ptr, on_device, int(t.device.device_id), shape, n * (t.dtype.bits // 8), cap, m)
        ok = True
    finally:
        # the capsule is consumed either way: a capsule renamed 'used_dltensor' is no longer freed by its destructor, so
        # the tensor is handed back to the producer here on failure and by Buffer.release() on success
        _api.PyCapsule_SetName(cap, _USED)
        if not ok and m.contents.deleter:
            m.contents.deleter(m)
    return buf


def resolve(a, what, dtypes=ANY, writable=False):
    """numpy array | object with __dlpack__ | DLPack capsule | legacy object with data_ptr() -> Buffer (or None)."""
    if a is None:
        return None
    if isinstance(a, Buffer):
        return a
    if isinstance(a, np.ndarray):
        return from_numpy(a, what, dtypes, writable)
    if hasattr(a, "__dlpack__") or type(a).__name__ == "PyCapsule":
        return from_dlpack(a, what, dtypes)
    if hasattr(a, "data_ptr"):  # tensor-like without DLPack: trusted only as far as it can be asked
        if hasattr(a, "is_contiguous") and not a.is_contiguous():
            _fail(what, "tensor must be contiguous")
        nbytes = int(a.numel()) * int(a.element_size()) if hasattr(a, "numel") and hasattr(a, "element_size") else 0
        if dtypes is not None and hasattr(a, "element_size") and all(a.element_size() * 8 != b for _, b in dtypes):
            _fail(what, "element size %d not accepted" % a.element_size())
        dev = getattr(getattr(a, "device", None), "index", None)
        return Buffer(int(a.data_ptr()), bool(getattr(a, "is_cuda", False)), -1 if dev is None else int(dev), getattr(a, "shape", ()), nbytes, a)
    raise TypeError("%s: expected a numpy array or a DLPack-capable buffer, got %r" % (what, type(a)))

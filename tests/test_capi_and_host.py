"""CPU: the C-ABI library loads and exports everything include/opp_b200.h declares (no compute calls
without a GPU), struct layouts match the reference's, and the host-side logic (result conversion,
sharding, gather over gloo with world_size 2)."""
import ctypes as C
import os
import re
import subprocess
import sys

import numpy as np
import pytest

import conftest
from openpose_plus_b200 import _capi as capi

ROOT = conftest.ROOT


def declared_functions():
    src = open(os.path.join(ROOT, "include", "opp_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = set(re.findall(r"\b(opp_[a-z_0-9]+|process_conf_paf)\s*\(", src))
    return sorted(names)


def test_library_exports_every_declared_symbol():
    L = capi.lib()
    names = declared_functions()
    assert len(names) >= 15
    for n in names:
        assert hasattr(L, n), n
    assert set(capi.EXPORTS) == set(names)
    assert b"sm_100a" in L.opp_version()


def test_struct_layouts():
    assert capi.HUMAN_DT.itemsize == 292 and capi.PEAK_DT.itemsize == 20 and capi.CONN_DT.itemsize == 12
    assert C.sizeof(capi.Config) == 16 * 4
    assert C.sizeof(capi.Batch) == 2 * 8 + 4 * 4 + 3 * 8 + 2 * 8 + 2 * 4 + 8
    assert capi.Batch.in_sync_obj.offset == 80 and capi.Batch.in_sync.offset == 76 and capi.Batch.up_layout.offset == 72


@pytest.mark.skipif(conftest.HAS_GPU, reason="only meaningful without a GPU")
def test_create_fails_loudly_without_gpu():
    from openpose_plus_b200.engine import Engine
    with pytest.raises(capi.OppError) as e:
        Engine(46, 54)
    assert e.value.code == capi.ERR_NO_DEVICE and "no CPU fallback" in str(e.value)


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "openpose_plus_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                for pat in (r"^\s*(from|import)\s+oracle", r"liborc", r"libopp_ref", r"#include\s*[<\"][^>\"]*oracle", r"orc_[a-z_]+\s*\("):
                    assert not re.search(pat, text, re.M), (os.path.join(dirpath, f), pat)


def test_cpp_dropin_header_compiles_and_links():
    """A caller written against the reference's API (create_paf_processor / paf_processor / human_t)
    builds against include/ and links with the library."""
    src = r'''
#include <memory>
#include <openpose-plus.h>
static_assert(sizeof(human_t) == 292, "human_t");
int main(int argc, char **) {
    if (argc > 100) {
        std::unique_ptr<paf_processor> p(create_paf_processor(46, 54, 368, 432, n_joins, n_connections, 17));
        std::vector<human_t> hs = (*p)(nullptr, nullptr, true);
        for (const auto &h : hs) h.print();
        process_conf_paf(46, 54, n_joins, n_connections, nullptr, nullptr);
    }
    return COCOPAIRS.size() == 19 && COCOPAIRS_NET.size() == 19 && is_virtual_pair(17) ? 0 : 1;
}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(cpp, "w").write(src)
        libdir = os.path.join(ROOT, "openpose_plus_b200")
        cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), cpp, "-o", exe, "-L", libdir, "-l:libopp_b200.so",
               "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert subprocess.run([exe]).returncode == 0


def test_humans_from_records_normalises_like_the_reference():
    from openpose_plus_b200.post_process import humans_from_records
    rec = np.zeros(2, capi.HUMAN_DT)
    rec[0]["score"] = 3.5
    rec[0]["parts"][1] = (1, [0, 0, 0], 216.0, 92.0, 0.9)
    rec[0]["parts"][4] = (1, [0, 0, 0], 0.0, 367.0, 0.5)
    hs = humans_from_records(rec, 368, 432)
    assert len(hs) == 1 and set(hs[0].body_parts) == {1, 4}
    bp = hs[0].body_parts[1]
    assert (bp.x, bp.y, bp.uidx, bp.part_idx) == (0.5, 0.25, "0-1", 1) and abs(bp.score - 0.9) < 1e-6
    assert hs[0].score == 3.5 and str(bp).startswith("BodyPart:1-(0.50, 0.25)")


def test_shard_ranges_partition_the_stream():
    from openpose_plus_b200.sharding import shard_range
    for n in (0, 1, 7, 64, 4096, 4097):
        for world in (1, 2, 3, 4, 8):
            r = [shard_range(n, k, world) for k in range(world)]
            assert r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in r]
            assert max(sizes) - min(sizes) <= 1


def test_synth_is_deterministic_and_in_range():
    from openpose_plus_b200 import synth
    a, b = synth.render_frame(42, 3), synth.render_frame(42, 3)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert a[0].shape == (19, 46, 54) and a[1].shape == (38, 46, 54)
    assert a[0].min() >= 0 and a[0].max() <= 1 and np.abs(a[1]).max() <= 1
    assert a[0][:18].max() > 0.5


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from openpose_plus_b200 import _capi as capi
from openpose_plus_b200.sharding import shard_range, gather_results
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")
n = 37
lo, hi = shard_range(n, rank, world)
humans = np.zeros((hi - lo, 4), capi.HUMAN_DT)
counts = (np.arange(lo, hi, dtype=np.int32) % 4).clip(1)     # only the humans found travel: slot 0 must be one of them
flags = np.zeros(hi - lo, np.int32)
for f in range(lo, hi):
    humans[f - lo]["score"] = f          # frame index rides in the payload
res = gather_results((humans, counts, flags), rank, world)
if rank == 0:
    H, Cn, F = res
    assert H.shape == (n, 4) and np.array_equal(H["score"][:, 0], np.arange(n, dtype=np.float32))
    assert np.array_equal(Cn, (np.arange(n, dtype=np.int32) % 4).clip(1))
    print("GATHER_OK")
else:
    assert res is None
dist.destroy_process_group()
'''


def test_host_gather_world_size_2_gloo(tmp_path):
    w = tmp_path / "worker.py"
    w.write_text(_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29731", str(w), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "GATHER_OK" in r.stdout, r.stdout + r.stderr


def test_draw_humans_marks_parts_and_limbs():
    from openpose_plus_b200.common import draw_humans, CocoColors
    from openpose_plus_b200.post_process import humans_from_records
    rec = np.zeros(1, capi.HUMAN_DT)
    rec[0]["parts"][1] = (1, [0, 0, 0], 100.0, 50.0, 0.9)
    rec[0]["parts"][2] = (1, [0, 0, 0], 150.0, 80.0, 0.9)
    rec[0]["parts"][10] = (1, [0, 0, 0], 300.0, 300.0, 0.9)
    img = draw_humans(np.zeros((368, 432, 3), np.uint8), humans_from_records(rec, 368, 432))
    assert tuple(img[300, 300]) == CocoColors[10]          # isolated part: its dot
    assert tuple(img[65, 125]) == CocoColors[0]            # midpoint of limb 0 (neck - right shoulder)
    assert img[200, 50].sum() == 0


def test_c_draw_human_overlay_and_cpp_header():
    """opp_draw_human (host code in the C-ABI library, the role of examples/vis.cpp:56-81): limb lines, part
    dots, clipping at the border, argument errors; and <openpose-plus/vis.h> compiles against it."""
    import ctypes as C
    L = capi.lib()
    rec = np.zeros(1, capi.HUMAN_DT)
    rec[0]["parts"][1] = (1, [0, 0, 0], 100.9, 50.2, 0.9)
    rec[0]["parts"][2] = (1, [0, 0, 0], 150.0, 80.0, 0.9)
    rec[0]["parts"][10] = (1, [0, 0, 0], 431.0, 367.0, 0.9)      # in the corner: must clip, not write outside
    guard = np.zeros((370, 432, 3), np.uint8)
    img = guard[1:369]
    assert L.opp_draw_human(img.ctypes.data, 368, 432, 3, 0, rec.ctypes.data, 2) == capi.OK
    from openpose_plus_b200.common import CocoColors
    assert tuple(img[367, 431]) == CocoColors[10] and tuple(img[80, 150]) == CocoColors[2] and tuple(img[50, 100]) == CocoColors[1]
    assert tuple(img[65, 125]) == CocoColors[0]                  # midpoint of limb 0
    assert img[200, 50].sum() == 0 and guard[0].sum() == 0 and guard[369].sum() == 0
    n_painted = int((img.sum(axis=2) > 0).sum())
    assert 150 < n_painted < 400                                 # a 2-px line of ~58 px and three small dots
    assert L.opp_draw_human(None, 368, 432, 3, 0, rec.ctypes.data, 2) == capi.ERR_INVALID
    assert L.opp_draw_human(img.ctypes.data, 368, 432, 5, 0, rec.ctypes.data, 2) == capi.ERR_INVALID
    src = r'''
#include <string>
#include <vector>
#include <openpose-plus.h>
#include <openpose-plus/vis.h>
int main() {
    std::vector<uint8_t> img(64 * 64 * 3, 0);
    human_t h;
    h.parts[0].has_value = true, h.parts[0].x = 10, h.parts[0].y = 12;
    h.parts[1].has_value = true, h.parts[1].x = 40, h.parts[1].y = 50;
    draw_human(img.data(), 64, 64, 3, h);
    return img[(12 * 64 + 10) * 3] == 255 && img[(31 * 64 + 25) * 3 + 2] == 255 ? 0 : 1;   // nose dot red, limb 12 blue
}
'''
    import tempfile
    with tempfile.TemporaryDirectory() as d:
        cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
        open(cpp, "w").write(src)
        libdir = os.path.join(ROOT, "openpose_plus_b200")
        cmd = ["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), cpp, "-o", exe, "-L", libdir, "-l:libopp_b200.so",
               "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        assert r.returncode == 0, r.stderr
        assert subprocess.run([exe]).returncode == 0


class _FakeEngine:
    """Stands in for Engine in host-logic tests: records what was submitted and writes frame indices as results."""

    def __init__(self, max_batch=8, n_slots=3, max_humans=4, fail_at=None):
        import types
        self.max_batch, self.max_humans, self.cfg = max_batch, max_humans, types.SimpleNamespace(n_slots=n_slots)
        self.calls, self.inflight, self.peak_inflight, self.fail_at = [], 0, 0, fail_at

    def submit(self, conf, paf, out=None, **kw):
        assert conf.shape[0] <= self.max_batch and self.inflight < self.cfg.n_slots
        if self.fail_at is not None and len(self.calls) == self.fail_at:
            raise capi.OppError(2, "injected")
        self.inflight += 1
        self.peak_inflight = max(self.peak_inflight, self.inflight)
        self.calls.append((conf, out))
        return len(self.calls) - 1

    def wait(self, t):
        conf, (humans, counts, flags) = self.calls[t]
        counts[:] = conf[:, 0, 0, 0].astype(np.int32)          # "number of humans" = the frame's tag
        humans["score"][:, 0] = conf[:, 0, 0, 0]
        flags[:] = 0
        self.inflight -= 1
        return humans, counts, flags


def test_stream_drivers_shard_order_and_pipeline_depth():
    """process_stream / process_stream_multi: contiguous shards, frame order kept, all slots used and never exceeded,
    ragged tails, more GPUs than frames, and an engine failure surfacing on the caller's thread."""
    from openpose_plus_b200.sharding import process_stream, process_stream_multi
    n = 53
    conf = np.zeros((n, 19, 2, 2), np.float32)
    conf[:, 0, 0, 0] = np.arange(n)
    paf = np.zeros((n, 38, 2, 2), np.float32)
    for world in (1, 2, 3, 8):
        engines = [_FakeEngine() for _ in range(world)]
        humans, counts, flags = process_stream_multi(engines, conf, paf)
        assert counts.tolist() == list(range(n)) and humans["score"][:, 0].tolist() == list(range(n))
        assert sum(sum(c[0].shape[0] for c in e.calls) for e in engines) == n
        assert all(e.inflight == 0 and e.peak_inflight <= 3 for e in engines)
        if n // world >= 3 * 8:                                 # enough batches per GPU to fill every slot
            assert all(e.peak_inflight == 3 for e in engines)
    e = _FakeEngine()
    h, c, f = process_stream(e, conf, paf, rank=0, world=1, batch=5)
    assert c.tolist() == list(range(n)) and [x[0].shape[0] for x in e.calls] == [5] * 10 + [3]
    few = process_stream_multi([_FakeEngine() for _ in range(8)], conf[:3], paf[:3])      # more GPUs than frames
    assert few[1].tolist() == [0, 1, 2]
    with pytest.raises(capi.OppError):
        process_stream_multi([_FakeEngine(), _FakeEngine(fail_at=1)], conf, paf)


def test_engine_takes_buffer_addresses_without_ndarray_ctypes():
    from openpose_plus_b200.engine import _ptr
    a = np.zeros((4, 5), np.float32)
    ro = a.view()
    ro.flags.writeable = False
    rec = np.zeros((2, 3), capi.HUMAN_DT)
    assert _ptr(a) == a.ctypes.data and _ptr(a[1:]) == a[1:].ctypes.data and _ptr(ro) == a.ctypes.data
    assert _ptr(rec[1:]) == rec[1:].ctypes.data and _ptr(None) is None
    with pytest.raises(TypeError):
        _ptr([1, 2, 3])


def test_create_rejects_bad_configurations_before_touching_the_device():
    """Argument errors are OPP_ERR_INVALID with a message, on any machine (the reference only asserts, src/post-process.h:37)."""
    L = capi.lib()

    def create(**kw):
        cfg = capi.Config()
        L.opp_config_default(C.byref(cfg), 46, 54, 368, 432, 17)
        for k, v in kw.items():
            setattr(cfg, k, v)
        h = C.c_void_p()
        rc = L.opp_create(C.byref(cfg), C.byref(h))
        if rc == capi.OK:
            L.opp_destroy(h)
        return rc, (L.opp_last_error(None) or b"").decode()

    for bad in (dict(n_joins=18), dict(n_connections=17), dict(gauss_kernel_size=16), dict(gauss_kernel_size=65), dict(gauss_kernel_size=0),
                dict(out_h=40), dict(feat_h=1), dict(out_w=40000), dict(variant=7), dict(max_peaks_per_part=100000), dict(max_humans=100000)):
        rc, msg = create(**bad)
        assert rc == capi.ERR_INVALID and msg.startswith("opp_create:"), (bad, rc, msg)
    rc, msg = create()
    assert rc == (capi.OK if conftest.HAS_GPU else capi.ERR_NO_DEVICE), (rc, msg)
    assert L.opp_bench_latency(None, None, 1, None) == capi.ERR_INVALID


REF_INC = "/root/reference/include"


@pytest.mark.skipif(not os.path.exists(os.path.join(REF_INC, "openpose-plus.hpp")), reason="reference headers only exist in the build container")
def test_dropin_builds_against_the_unmodified_reference_headers(tmp_path):
    """The ABI claim pinned in CI: a caller compiled with the reference's OWN headers links against libopp_b200.so
    (factory signature, paf_processor vtable), and the structs those headers define have the layout the C-ABI writes."""
    from openpose_plus_b200 import build as b
    exe = b.build_dropin_against_reference_headers(force=True)
    assert exe and os.path.exists(exe)
    nm = subprocess.run(["nm", "-D", "--undefined-only", exe], capture_output=True, text=True).stdout
    assert "create_paf_processor" in nm
    probe = tmp_path / "abi.cpp"
    probe.write_text(r"""
#include <cstddef>
#include <string>
#include <openpose-plus.hpp>          // the reference's
#include "opp_b200.h"                 // ours
static_assert(sizeof(human_t) == sizeof(opp_human_t) && sizeof(human_t) == 292, "human_t");
static_assert(sizeof(body_part_t) == sizeof(opp_body_part_t), "body_part_t");
static_assert(offsetof(body_part_t, x) == offsetof(opp_body_part_t, x) && offsetof(body_part_t, score) == offsetof(opp_body_part_t, score), "part fields");
static_assert(offsetof(human_t, score) == offsetof(opp_human_t, score), "human score");
// vtable order of the interface the library implements: operator() first, then the destructors
struct probe_impl : paf_processor {
    std::vector<human_t> operator()(const float *, const float *, bool) override { return {}; }
};
int main() { probe_impl p; paf_processor *q = &p; return (int)(*q)(nullptr, nullptr, false).size(); }
""")
    r = subprocess.run(["g++", "-std=c++14", "-I", REF_INC, "-I", os.path.join(ROOT, "include"), str(probe), "-o", str(tmp_path / "abi")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    # the re-written headers this repo ships declare the same interface (incl. the out-of-scope runner, declaration only)
    ours = open(os.path.join(ROOT, "include", "openpose-plus.hpp")).read()
    assert "class pose_detection_runner" in ours and "create_pose_detection_runner(" in ours


def test_shipped_header_lets_a_runner_translation_unit_compile(tmp_path):
    """with this repo's include/ first on the path, code that implements or calls the reference's
    pose_detection_runner (src/uff-runner.cpp, examples/*.cpp) still compiles."""
    src = tmp_path / "runner.cpp"
    src.write_text(r"""
#include <openpose-plus.h>
struct fake_runner : pose_detection_runner {
    void operator()(const std::vector<void *> &, const std::vector<void *> &, int) override {}
};
pose_detection_runner *create_pose_detection_runner(const std::string &, int, int, int, bool) { return new fake_runner; }
int main() { delete create_pose_detection_runner("m.uff", 368, 432, 1, false); return 0; }
""")
    r = subprocess.run(["g++", "-std=c++14", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(tmp_path / "runner")], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    assert subprocess.run([str(tmp_path / "runner")]).returncode == 0


def test_tranform_keypoints2d_matches_the_reference_helper():
    """openpose_plus/inference/common.py:86-96 (name sic): pixel coordinates, scores, visibility above the threshold."""
    from openpose_plus_b200.common import tranform_keypoints2d
    from openpose_plus_b200.post_process import BodyPart
    body = {0: BodyPart("0-0", 0, 0.5, 0.25, 0.9), 7: BodyPart("0-7", 7, 0.1, 0.75, 0.2), 17: BodyPart("0-17", 17, 1.0, 1.0, 0.25)}
    xy, conf, vis = tranform_keypoints2d(body, 432, 368)
    assert xy.shape == (18, 2) and conf.shape == (18,) and vis.dtype == bool
    assert xy[0].tolist() == [216.0, 92.0] and xy[7].tolist() == [0.1 * 432, 0.75 * 368] and xy[3].tolist() == [0.0, 0.0]
    assert conf[0] == 0.9 and conf[7] == 0.2 and conf[1] == 0
    assert vis.tolist() == [i == 0 for i in range(18)]          # 0.25 is not above the default threshold
    assert tranform_keypoints2d(body, 432, 368, kp_score_thresh=0.1)[2].sum() == 3


def test_dlpack_intake_validates_and_consumes():
    """north_star: 'numpy or DLPack buffers'.  Capsules are consumed through the CPython capsule API (no torch import
    inside the package), validated (device / dtype / compact strides) and handed back to the producer on release."""
    import gc
    import torch
    from openpose_plus_b200 import dlpack
    t = torch.arange(2 * 19 * 4 * 6, dtype=torch.float32).reshape(2, 19, 4, 6)
    b = dlpack.resolve(t, "conf", dlpack.F32)
    assert b.ptr == t.data_ptr() and b.shape == (2, 19, 4, 6) and b.nbytes == t.numel() * 4 and not b.on_device
    b.release()
    b.release()                                                       # idempotent
    sl = dlpack.resolve(t[1:], "conf", dlpack.F32)                    # contiguous slice: byte offset / data pointer honoured
    assert sl.ptr == t[1:].data_ptr() and sl.shape == (1, 19, 4, 6)
    cap = t.__dlpack__()                                              # a ready capsule (what a C++ producer would hand over)
    assert dlpack.resolve(cap, "conf", dlpack.F32).ptr == t.data_ptr()
    with pytest.raises(TypeError):
        dlpack.resolve(cap, "conf", dlpack.F32)                       # already consumed
    a = np.zeros((3, 5), np.float32)
    assert dlpack.from_dlpack(a, "a", dlpack.F32).ptr == a.ctypes.data  # numpy speaks DLPack too
    for bad in (t[:, :, ::2], t.transpose(2, 3), t.double(), t.half(), t.int()):
        with pytest.raises(TypeError):
            dlpack.resolve(bad, "conf", dlpack.F32)
    with pytest.raises(TypeError):
        dlpack.resolve(np.zeros((4, 4), np.float32)[:, ::2], "x", dlpack.F32)
    with pytest.raises(TypeError):
        dlpack.resolve("not a buffer", "x")
    assert dlpack.resolve(torch.zeros(4, dtype=torch.int32), "n", dlpack.I32).nbytes == 16
    # the producer gets its tensor back: nothing keeps the storage alive after release
    import weakref
    u = torch.zeros(1000)
    ref = weakref.ref(u.untyped_storage())
    bu = dlpack.resolve(u, "u")
    del u
    gc.collect()
    assert ref() is not None                                          # the consumed capsule owns a reference
    bu.release()
    del bu
    gc.collect()
    assert ref() is None


def test_engine_rejects_wrong_shapes_dtypes_and_strides_before_the_c_abi():
    """The C-ABI takes raw pointers: the Python host side refuses anything it cannot vouch for (a non-contiguous
    or fp16 device tensor gave wrong skeletons silently)."""
    import torch
    from openpose_plus_b200.engine import Engine
    e = object.__new__(Engine)                                        # host-side checks only: no device, no handle
    e.feat, e.out, e.max_humans, e.device = (4, 6), (32, 48), 8, 0
    ok = np.zeros((2, 19, 4, 6), np.float32)
    a, ptr, n, dev = e._map_in(ok, 19, "conf", capi.LAYOUT_CHW)
    assert ptr == ok.ctypes.data and n == 2 and not dev
    a, ptr, n, dev = e._map_in(ok.astype(np.float64)[:, :, ::1], 19, "conf", capi.LAYOUT_CHW)   # numpy inputs are normalised
    assert a.dtype == np.float32 and ptr == a.ctypes.data
    assert e._map_in(np.zeros((2, 4, 6, 38), np.float32), 38, "paf", capi.LAYOUT_HWC)[2] == 2
    for bad in (np.zeros((2, 19, 4, 5), np.float32), np.zeros((19, 4, 6), np.float32), np.zeros((2, 38, 4, 6), np.float32)):
        with pytest.raises(capi.OppError):
            e._map_in(bad, 19, "conf", capi.LAYOUT_CHW)
    t = torch.zeros(2, 19, 4, 6)
    assert e._map_in(t, 19, "conf", capi.LAYOUT_CHW)[1] == t.data_ptr()
    for bad in (t.half(), t.double(), torch.zeros(2, 19, 4, 12)[..., ::2], torch.zeros(2, 19, 6, 4).transpose(2, 3)):
        with pytest.raises(capi.OppError):
            e._map_in(bad, 19, "conf", capi.LAYOUT_CHW)
    humans = np.zeros((2, 8), capi.HUMAN_DT)
    assert e._buf_out(humans, "humans", 2 * 8 * 292)[1] == humans.ctypes.data
    with pytest.raises(capi.OppError):
        e._buf_out(humans[:1], "humans", 2 * 8 * 292)                 # too small
    with pytest.raises(capi.OppError):
        e._buf_out(np.zeros(2, np.int64), "n_humans", 8, 4)           # wrong element size
    with pytest.raises(capi.OppError):
        e._buf_out(np.zeros((2, 19, 32, 48), np.float32), "conf_up", 4, device_only=True)   # up-sampled maps are device buffers
    ro = np.zeros(2, np.int32)
    ro.flags.writeable = False
    with pytest.raises(capi.OppError):
        e._buf_out(ro, "n_humans", 8, 4)


_SHM_WORKER = r"""
import os, sys, time
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
sys.path.insert(0, os.path.join(sys.argv[1], "tests"))
from openpose_plus_b200.sharding import HostGather, process_stream, shard_range
from test_capi_and_host import _FakeEngine
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo")                       # only for the final cross-check; the gather itself is shared memory
n = 53
conf = np.zeros((n, 19, 2, 2), np.float32)
conf[:, 0, 0, 0] = np.arange(n) % 4                   # "humans found" per frame rides in the payload
paf = np.zeros((n, 38, 2, 2), np.float32)
g = HostGather("opp_test_gather_%s" % os.environ["MASTER_PORT"], n, 4, rank, world, register=False)
lo, hi = shard_range(n, rank, world)
for rnd in range(3):                                  # three passes over the stream through the same segment
    if rnd:
        g.wait_released() if rank else None
    eng = _FakeEngine()
    if rnd == 1:                                      # a rank that only HOLDS its shard
        res = process_stream(eng, conf[lo:hi], paf[lo:hi], rank, world, gather=g, shard_only=True)
    else:
        res = process_stream(eng, conf, paf, rank, world, gather=g)
    assert sum(c[0].shape[0] for c in eng.calls) == hi - lo
    if rank == 0:
        H, Cn, F = res
        assert Cn.tolist() == (np.arange(n) % 4).tolist() and H["score"][:, 0].tolist() == (np.arange(n) % 4).astype(float).tolist()
        H["score"][:] = -1                            # consumed; the next round must rewrite every frame
        Cn[:] = -1
        g.release()
    else:
        assert res is None
dist.barrier()
g.close()
if rank == 0:
    assert not os.path.exists("/dev/shm/opp_test_gather_%s" % os.environ["MASTER_PORT"])
    print("SHM_GATHER_OK")
dist.destroy_process_group()
"""


def test_shared_memory_host_gather_world_size_2(tmp_path):
    """configs[4]'s host gather as bench.py runs it: every rank writes its shard's results into one shared segment and
    rank 0 sees the stream in frame order after a sequence-number handshake - no copy, no collective."""
    w = tmp_path / "worker.py"
    w.write_text(_SHM_WORKER)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29741", str(w), ROOT]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert r.returncode == 0 and "SHM_GATHER_OK" in r.stdout, r.stdout + r.stderr


def test_gather_results_ships_only_the_humans_found():
    from openpose_plus_b200.sharding import _compact, _expand
    hh = np.zeros((5, 4), capi.HUMAN_DT)
    hh["score"] = np.arange(20).reshape(5, 4)
    cc = np.array([0, 1, 4, 2, 9], np.int32)              # 9 > capacity: an overflowed frame keeps its 4 slots
    recs, c2, f2, cap = _compact((hh, cc, cc))
    assert len(recs) == 0 + 1 + 4 + 2 + 4 and cap == 4
    back = _expand((recs, c2, f2, cap))
    want = hh.copy()
    for f in range(5):
        want["score"][f, min(cc[f], 4):] = 0
    assert np.array_equal(back[0]["score"], want["score"]) and np.array_equal(back[1], cc)
    empty = _expand(_compact((np.zeros((0, 4), capi.HUMAN_DT), np.zeros(0, np.int32), np.zeros(0, np.int32))))
    assert empty[0].shape == (0, 4)


def test_no_fused_multiply_add_in_the_filter_kernels():
    """The resize / peak kernels must round every product before adding it (bit-exactness with cv::GaussianBlur's scalar
    path): no FFMA / FFMA2 in their SASS - in particular none contracted from the packed adds' operands."""
    from openpose_plus_b200 import build
    if not os.path.exists(build.CUOBJDUMP):
        pytest.skip("cuobjdump not available")
    assert build.fused_multiply_adds() == {}
    sass = subprocess.run([build.CUOBJDUMP, "-sass", build.OUT], check=True, capture_output=True, text=True).stdout
    assert "FADD2" in sass  # the packed adds are really there

"""-m gpu: the drop-in boundary beyond plain host arrays - DLPack / device buffers with validation, producer-stream
ordering for device-resident maps (N2), the reference-header C++ drop-in, process_conf_paf, capacity growth behind the
reference-shaped entries, write-combined and registered host memory, and geometries at the limits of the tile planner."""
import os
import subprocess
import sys

import numpy as np
import pytest

import conftest
from openpose_plus_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
ROOT = conftest.ROOT


@pytest.fixture(scope="module")
def mods():
    from openpose_plus_b200.engine import Engine
    from openpose_plus_b200 import _capi as capi
    from oracle.oracle import Oracle
    import helpers
    return Engine, capi, helpers, Oracle


def _same(H, humans, counts, ref_h, ref_c):
    assert np.array_equal(counts, ref_c)
    for f in range(len(counts)):
        assert H.humans_equal(humans[f, :counts[f]], ref_h[f, :ref_c[f]]) is None, f


def test_dlpack_device_buffers_in_and_out(mods):
    """Any __dlpack__ object is accepted (north_star: 'numpy or DLPack buffers'): CUDA tensors, host tensors, ready
    capsules; results may stay on the device.  Wrong dtype / strides / device are errors, not wrong skeletons."""
    import torch
    Engine, capi, H, Oracle = mods
    conf, paf = synth.render_batch(4, n_people=5, seed0=910)
    eng = Engine(46, 54, max_batch=4)
    ref_h, ref_c, _ = eng.process(conf, paf)
    dc, dp = torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda()

    class OnlyDlpack:                                     # exposes nothing but the protocol (what cupy / jax arrays look like to us)
        def __init__(self, t):
            self.t = t

        def __dlpack__(self, **kw):
            return self.t.__dlpack__()

        def __dlpack_device__(self):
            return self.t.__dlpack_device__()

    for c, p in ((dc, dp), (OnlyDlpack(dc), OnlyDlpack(dp)), (dc.__dlpack__(), dp.__dlpack__()), (torch.from_numpy(conf), torch.from_numpy(paf))):
        h, n, f = eng.process(c, p)
        _same(H, h, n, ref_h, ref_c)
    # results on the device, as DLPack buffers
    dh = torch.zeros((4, eng.max_humans, 292), dtype=torch.uint8, device="cuda")
    dn, df = torch.zeros(4, dtype=torch.int32, device="cuda"), torch.zeros(4, dtype=torch.int32, device="cuda")
    eng.process(dc, dp, out=(dh, dn, df))
    _same(H, dh.cpu().numpy().view(capi.HUMAN_DT).reshape(4, eng.max_humans), dn.cpu().numpy(), ref_h, ref_c)
    # channels-last device maps that are only a VIEW (non-compact) must be refused, as must fp16 / fp64
    for bad_c, bad_p in ((dc.permute(0, 2, 3, 1), dp.permute(0, 2, 3, 1)), (dc.half(), dp.half()), (dc.double(), dp.double()),
                         (dc[:, :, :, ::2], dp[:, :, :, ::2])):
        with pytest.raises(capi.OppError):
            eng.process(bad_c, bad_p, layout=capi.LAYOUT_HWC if bad_c.shape[-1] == 19 else capi.LAYOUT_CHW)
    with pytest.raises(capi.OppError):
        eng.process(dc, dp, conf_up=torch.empty((4, 19, 368, 432)), paf_up=torch.empty((4, 38, 368, 432)))   # host tensors for device outputs
    with pytest.raises(capi.OppError):
        eng.process(dc, dp, conf_up=torch.empty((2, 19, 368, 432), device="cuda"), paf_up=torch.empty((4, 38, 368, 432), device="cuda"))  # too small
    h, n, f = eng.process(dc.permute(0, 2, 3, 1).contiguous(), dp.permute(0, 2, 3, 1).contiguous(), layout=capi.LAYOUT_HWC)
    _same(H, h, n, ref_h, ref_c)


def test_producer_stream_ordering_without_host_sync(mods):
    """N2 (src/uff-runner.cpp:199-217 -> include/openpose-plus.hpp:42-51): the maps are written by a producer on ITS
    stream; with in_stream / in_event the hand-off is ordered on the device and the host never synchronises."""
    import torch
    Engine, capi, H, Oracle = mods
    conf, paf = synth.render_batch(8, n_people=5, seed0=930)
    eng = Engine(46, 54, max_batch=8)
    ref_h, ref_c, _ = eng.process(conf, paf)
    assert ref_c.sum() > 0
    src_c, src_p = torch.from_numpy(conf).cuda(), torch.from_numpy(paf).cuda()
    dc, dp = torch.zeros_like(src_c), torch.zeros_like(src_p)
    prod = torch.cuda.Stream()
    torch.cuda.synchronize()

    def produce():
        dc.zero_(), dp.zero_()
        torch.cuda.synchronize()
        with torch.cuda.stream(prod):
            torch.cuda._sleep(60_000_000)                  # the "CNN": tens of milliseconds before the maps exist
            dc.copy_(src_c, non_blocking=True)
            dp.copy_(src_p, non_blocking=True)

    produce()
    h, n, f = eng.process(dc, dp, in_stream=prod)          # no host synchronisation between producer and submit
    _same(H, h, n, ref_h, ref_c)
    produce()
    ev = torch.cuda.Event()
    ev.record(prod)
    h, n, f = eng.process(dc, dp, in_event=ev)
    _same(H, h, n, ref_h, ref_c)
    # control: without the ordering the batch reads the maps before the producer wrote them (all zeros -> nobody)
    produce()
    h, n, f = eng.process(dc, dp)
    assert n.sum() == 0
    torch.cuda.synchronize()
    # consumer side: a stream of the caller waits for the ticket on the device (results stay on the device)
    dh = torch.zeros((8, eng.max_humans, 292), dtype=torch.uint8, device="cuda")
    dn, df = torch.zeros(8, dtype=torch.int32, device="cuda"), torch.zeros(8, dtype=torch.int32, device="cuda")
    cons = torch.cuda.Stream()
    t = eng.submit(src_c, src_p, out=(dh, dn, df))
    eng.stream_wait(t, cons)
    with torch.cuda.stream(cons):
        got = dn.clone()
    eng.wait(t)
    cons.synchronize()
    assert np.array_equal(got.cpu().numpy(), ref_c)
    with pytest.raises(capi.OppError):
        eng.stream_wait(12345, cons)


def _run_dropin(exe, capi, H, tmp_path):
    for name in ("frame_5p_368x432_k17", "frame_35p_368x432_k13", "frame_6p_300x400_k17"):
        g = np.load(os.path.join(GOLD, name + ".npz"))
        fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "out.bin")
        with open(fin, "wb") as f:
            f.write(g["geom"].astype(np.int32).tobytes())
            f.write(np.int32(2).tobytes())
            for _ in range(2):
                f.write(np.ascontiguousarray(g["conf"], np.float32).tobytes())
                f.write(np.ascontiguousarray(g["paf"], np.float32).tobytes())
        r = subprocess.run([exe, fin, fout], capture_output=True, text=True, timeout=120)
        assert r.returncode == 0, r.stdout + r.stderr
        raw = open(fout, "rb").read()
        off = 0
        for _ in range(2):
            m = int(np.frombuffer(raw, np.int32, 1, off)[0])
            off += 4
            humans = np.frombuffer(raw, capi.HUMAN_DT, m, off)
            off += m * 292
            assert H.humans_equal(humans, g["humans_ref"]) is None, name
        assert off == len(raw)


def test_cpp_dropin_compiled_with_the_reference_headers(mods, tmp_path):
    """tests/cpp/dropin_main.cpp built against the reference's UNMODIFIED include/ (by build(), in the container that has
    the reference tree; the binary travels) and linked with libopp_b200.so reproduces the reference build's golden humans."""
    from openpose_plus_b200 import build as b
    Engine, capi, H, Oracle = mods
    exe = b.build_dropin_against_reference_headers()
    if exe is None:
        pytest.skip("no reference headers here and no prebuilt tests/cpp/_build/dropin_refhdr (run __graft_entry__.build() where /root/reference exists)")
    _run_dropin(exe, capi, H, tmp_path)


_PCP = r'''
import ctypes as C, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from openpose_plus_b200 import _capi as capi
g = np.load(sys.argv[2])
conf, paf = np.ascontiguousarray(g["conf"], np.float32), np.ascontiguousarray(g["paf"], np.float32)
L = capi.lib()
L.process_conf_paf.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_void_p]
L.process_conf_paf.restype = None
L.process_conf_paf(conf.shape[1], conf.shape[2], 19, 19, conf.ctypes.data, paf.ctypes.data)
C.CDLL(None).fflush(None)
'''


def test_process_conf_paf_prints_the_golden_humans(mods, tmp_path):
    """include/openpose-plus.h:18-22: declared by the reference, defined by this library.  One golden frame through it;
    its stdout (human_t::print format, include/openpose-plus/human.h:21-31) must name exactly the golden humans."""
    Engine, capi, H, Oracle = mods
    path = os.path.join(GOLD, "frame_5p_368x432_k17.npz")
    g = np.load(path)
    w = tmp_path / "pcp.py"
    w.write_text(_PCP)
    r = subprocess.run([sys.executable, str(w), ROOT, path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    want = []
    for hu in g["humans_ref"]:
        s = ""
        for j in range(18):
            p = hu["parts"][j]
            if p["has_value"]:
                s += "BodyPart:%d-(%.2f, %.2f) score=%.2f " % (j, p["x"], p["y"], p["score"])
        want.append(s + "score=%.2f" % hu["score"])
    got = [ln for ln in r.stdout.splitlines() if "score=" in ln]
    assert len(want) >= 3 and got == want, (got, want)


def test_reference_shaped_entries_grow_capacities_instead_of_truncating(mods):
    """The reference's vectors are unbounded (src/post-process.h:190-198); the drop-in entries must never hand back a
    truncated frame: PostProcessor re-creates its engine with larger capacities when a frame overflows them."""
    from openpose_plus_b200 import PostProcessor
    Engine, capi, H, Oracle = mods
    fr = synth.render_frame(5207, n_people=34)
    conf, paf = fr
    o = Oracle(46, 54, 368, 432, 17).run(conf, paf, lazy=True)
    assert len(o["peaks"]) > 18 * 8
    pp = PostProcessor((368, 432), (46, 54), "channels_first", return_maps=False, max_peaks_per_part=8, max_cands_per_limb=16, max_humans=4)
    humans, _, _ = pp(conf, paf)
    assert pp.engine.cfg.max_peaks_per_part > 8 and pp.engine.cfg.max_humans > 4
    assert len(humans) == o["n_humans"]
    for hu, rec in zip(humans, o["humans"]):
        assert np.float32(hu.score) == rec["score"]
        assert sorted(hu.body_parts) == [j for j in range(18) if rec["parts"][j]["has_value"]]
    # the raw engine keeps reporting instead (callers of the C-ABI see the flags)
    eng = Engine(46, 54, max_batch=1, max_peaks_per_part=8)
    assert eng.process(conf[None], paf[None])[2][0] & capi.FLAG_PEAK_OVERFLOW


def test_write_combined_and_registered_host_memory(mods):
    """Input rings in write-combined pinned memory (the host only writes them) and result buffers in memory the caller
    owns and registers (e.g. a shared-memory segment all ranks of a box map: the host gather without a copy)."""
    import mmap
    import ctypes as C
    Engine, capi, H, Oracle = mods
    conf, paf = synth.render_batch(3, n_people=5, seed0=950)
    eng = Engine(46, 54, max_batch=3)
    ref_h, ref_c, _ = eng.process(conf, paf)
    wc_c, wc_p = capi.pinned_empty(conf.shape, np.float32, write_combined=True), capi.pinned_empty(paf.shape, np.float32, write_combined=True)
    wc_c[...] = conf
    wc_p[...] = paf
    for n in (1, 3):                                       # latency path (read in place by the kernels) and copy path
        h, c, f = eng.process(wc_c[:n], wc_p[:n])
        _same(H, h, c, ref_h[:n], ref_c[:n])
    nbytes = 3 * eng.max_humans * 292 + 2 * 3 * 4
    size = (nbytes + mmap.PAGESIZE - 1) // mmap.PAGESIZE * mmap.PAGESIZE
    mm = mmap.mmap(-1, size)
    base = C.addressof(C.c_char.from_buffer(mm))
    L = capi.lib()
    assert L.opp_host_register(base, size) == capi.OK
    try:
        buf = np.frombuffer(mm, np.uint8)
        hh = buf[:3 * eng.max_humans * 292].view(capi.HUMAN_DT).reshape(3, eng.max_humans)
        cc = buf[3 * eng.max_humans * 292:][:12].view(np.int32)
        ff = buf[3 * eng.max_humans * 292 + 12:][:12].view(np.int32)
        eng.process(conf, paf, out=(hh, cc, ff))           # written by the assembly kernel straight into the registered pages
        _same(H, hh, cc, ref_h, ref_c)
        del hh, cc, ff, buf
    finally:
        assert L.opp_host_unregister(base) == capi.OK
        mm.close()


@pytest.mark.parametrize("geom", [
    (46, 54, 4, 9, "x4, fused store of both maps (ADVICE r1: every x4 STORE launch was rejected)"),
    (20, 70, 4, 7, "x4, feature width 63..108: one column strip would exceed 62 columns"),
    (12, 130, 4, 9, "x4, width 125+"),
    (10, 217, 8, 17, "x8, width 217: 55-column strips would need 8 column groups"),
    (9, 271, 8, 13, "x8, width 271"),
    (70, 30, 8, 17, "x8, 70 feature rows: more than one 60-row tile"),
], ids=lambda g: "%dx%d_x%d_k%d" % g[:4])
def test_tile_planner_limits(mods, geom):
    """Geometries at the limits of the integer-scale peak kernel's tile plan: peaks, skeletons and both materialised maps
    (requested channels-first, i.e. the fused-store variant where it applies) against the oracle."""
    import torch
    Engine, capi, H, Oracle = mods
    fh, fw, S, k, _ = geom
    oh, ow = fh * S, fw * S
    rng = np.random.default_rng(fh * 1000 + fw)
    n = 2
    conf, paf = synth.render_batch(n, n_people=4, feat_h=fh, feat_w=fw, stride=S, seed0=7000)
    conf = np.clip(conf + rng.uniform(0, 0.08, conf.shape).astype(np.float32), 0, 1)   # every block active, extra peaks at the strip seams
    eng, orc = Engine(fh, fw, oh, ow, k, max_batch=n, max_peaks_per_part=1024, max_cands_per_limb=16384, max_humans=512), Oracle(fh, fw, oh, ow, k)
    cu, pu = torch.empty((n, 19, oh, ow), device="cuda"), torch.empty((n, 38, oh, ow), device="cuda")
    for kw in ({}, dict(conf_up=cu, paf_up=pu)):
        t = eng.submit(conf, paf, **kw)
        humans, counts, flags = eng.wait(t)
        for f in range(n):
            o = orc.run(conf[f], paf[f], maps=bool(kw))
            H.check_frame(eng, t, f, humans, counts, flags, o, geom[4])
            if kw:
                assert np.array_equal(cu[f].cpu().numpy(), o["conf_up"]) and np.array_equal(pu[f].cpu().numpy(), o["paf_up"])

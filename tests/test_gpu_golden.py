"""-m gpu: the CUDA path against the committed golden vectors (reference build outputs), the
batched / pipelined / sharded host drivers, materialised maps, channels-last input, and
size-independent properties at BASELINE.json's full sizes."""
import glob
import os

import numpy as np
import pytest

from openpose_plus_b200 import synth

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")
FRAMES = sorted(glob.glob(os.path.join(GOLD, "frame_*.npz")))


@pytest.fixture(scope="module")
def mods():
    from openpose_plus_b200.engine import Engine
    from openpose_plus_b200 import _capi as capi
    import helpers
    return Engine, capi, helpers


@pytest.mark.parametrize("path", FRAMES, ids=[os.path.basename(p)[:-4] for p in FRAMES])
def test_cuda_reproduces_reference_golden(mods, path):
    Engine, capi, H = mods
    g = np.load(path)
    h, w, oh, ow, k = [int(v) for v in g["geom"]]
    eng = Engine(h, w, oh, ow, k, max_batch=1, max_peaks_per_part=512, max_cands_per_limb=8192, max_humans=512)
    t = eng.submit(g["conf"][None], g["paf"][None])
    humans, counts, flags = eng.wait(t)
    assert (flags[0] & capi.FLAG_OVERFLOW_MASK) == 0
    assert np.array_equal(eng.debug_peaks(t, 0, cap=18 * 512).view(np.uint8), g["peaks"].view(np.uint8))
    for p in range(19):
        assert np.array_equal(eng.debug_conns(t, 0, p).view(np.uint8), g["conns_%02d" % p].view(np.uint8)), p
    assert H.humans_equal(humans[0, :counts[0]], g["humans_ref"]) is None
    for i in range(counts[0]):
        assert np.array_equal(eng.debug_parts(t, 0, i), g["hrefs"]["parts"][i])


def test_cuda_reproduces_python_variant_vectors(mods):
    """OPP_VARIANT_PYTHON against its committed regression vectors (oracle variant 1; unpinned, see
    scripts/make_golden_python_variant.py): peaks, connections, humans and person -> part ids, byte for byte."""
    Engine, capi, H = mods
    paths = sorted(glob.glob(os.path.join(GOLD, "pyvariant_*.npz")))
    assert len(paths) >= 2
    for path in paths:
        g = np.load(path)
        h, w, oh, ow, k = [int(v) for v in g["geom"]]
        eng = Engine(h, w, oh, ow, k, max_batch=1, max_peaks_per_part=512, max_cands_per_limb=8192, max_humans=512, variant=capi.VARIANT_PYTHON)
        t = eng.submit(g["conf"][None], g["paf"][None])
        humans, counts, flags = eng.wait(t)
        assert flags[0] == 0
        assert np.array_equal(eng.debug_peaks(t, 0, cap=18 * 512).view(np.uint8), g["peaks"].view(np.uint8))
        for p in range(19):
            assert np.array_equal(eng.debug_conns(t, 0, p).view(np.uint8), g["conns_%02d" % p].view(np.uint8)), p
        assert H.humans_equal(humans[0, :counts[0]], g["humans"]) is None
        for i in range(counts[0]):
            assert np.array_equal(eng.debug_parts(t, 0, i), g["hrefs"]["parts"][i])
        eng.close()


def test_materialised_maps_and_layouts(mods):
    import torch
    from oracle.oracle import Oracle
    Engine, capi, H = mods
    conf, paf = synth.render_batch(3, n_people=4, seed0=70)
    for (oh, ow) in [(368, 432), (300, 400)]:
        eng, orc = Engine(46, 54, oh, ow, 17, max_batch=3), Oracle(46, 54, oh, ow, 17)
        cu = torch.empty((3, 19, oh, ow), device="cuda")
        pu = torch.empty((3, 38, oh, ow), device="cuda")
        humans, counts, flags = eng.process(conf, paf, conf_up=cu, paf_up=pu)
        cu2 = torch.empty((3, oh, ow, 19), device="cuda")
        pu2 = torch.empty((3, oh, ow, 38), device="cuda")
        # channels-last in, channels-last maps out (the Python PostProcessor contract)
        h2, c2, f2 = eng.process(np.ascontiguousarray(conf.transpose(0, 2, 3, 1)), np.ascontiguousarray(paf.transpose(0, 2, 3, 1)),
                                 layout=capi.LAYOUT_HWC, conf_up=cu2, paf_up=pu2, up_layout=capi.LAYOUT_HWC)
        assert np.array_equal(counts, c2)
        for f in range(3):
            o = orc.run(conf[f], paf[f], maps=True)
            assert np.array_equal(cu[f].cpu().numpy(), o["conf_up"]) and np.array_equal(pu[f].cpu().numpy(), o["paf_up"])
            assert np.array_equal(cu2[f].cpu().numpy(), o["conf_up"].transpose(1, 2, 0))
            assert np.array_equal(pu2[f].cpu().numpy(), o["paf_up"].transpose(1, 2, 0))
            assert H.humans_equal(humans[f, :counts[f]], o["humans"]) is None
            assert H.humans_equal(h2[f, :c2[f]], o["humans"]) is None


def test_device_resident_inputs_and_pipelined_stream(mods):
    """4096-frame stream (BASELINE configs[4]) through submit/wait with all slots in flight, device
    inputs; properties: every frame equals the result of its source frame processed alone."""
    import torch
    from openpose_plus_b200.sharding import process_stream
    Engine, capi, H = mods
    base_c, base_p = synth.render_batch(16, n_people=5, seed0=500)
    idx = np.arange(4096) % 16
    conf = torch.from_numpy(base_c).cuda()[torch.from_numpy(idx).cuda()]
    paf = torch.from_numpy(base_p).cuda()[torch.from_numpy(idx).cuda()]
    eng = Engine(46, 54, max_batch=64)
    humans, counts, flags = process_stream(eng, conf, paf)
    ref_h, ref_c, ref_f = eng.process(base_c, base_p)
    assert not (flags & capi.FLAG_OVERFLOW_MASK).any()
    assert np.array_equal(counts, ref_c[idx]) and np.array_equal(flags, ref_f[idx])
    for f in range(0, 4096, 97):
        assert H.humans_equal(humans[f, :counts[f]], ref_h[idx[f], :ref_c[idx[f]]]) is None


def test_python_postprocessor_dropin(mods):
    from oracle.oracle import Oracle
    from openpose_plus_b200 import PostProcessor
    conf, paf = synth.render_frame(77, 4)
    orc = Oracle(46, 54, 368, 432, 17).run(conf, paf, maps=True)
    for fmt in ("channels_first", "channels_last"):
        pp = PostProcessor((368, 432), (46, 54), fmt)
        hm, pm = (conf, paf) if fmt == "channels_first" else (conf.transpose(1, 2, 0), paf.transpose(1, 2, 0))
        humans, hup, pup = pp(hm, pm)
        assert hup.shape == (368, 432, 19) and pup.shape == (368, 432, 38)
        assert np.array_equal(hup, orc["conf_up"].transpose(1, 2, 0)) and np.array_equal(pup, orc["paf_up"].transpose(1, 2, 0))
        assert len(humans) == orc["n_humans"]
        for hu, rec in zip(humans, orc["humans"]):
            assert abs(hu.score - rec["score"]) <= 1e-5 * abs(rec["score"])
            for idx, bp in hu.body_parts.items():
                assert rec["parts"][idx]["has_value"]
                assert bp.x == rec["parts"][idx]["x"] / 432 and bp.y == rec["parts"][idx]["y"] / 368


def test_capacity_overflow_is_flagged_not_silent(mods):
    Engine, capi, H = mods
    conf, paf = synth.noise_frame(3)
    eng = Engine(46, 54, max_batch=1)
    humans, counts, flags = eng.process(conf[None], paf[None])
    assert flags[0] & capi.FLAG_PEAK_OVERFLOW


def test_bad_arguments_are_errors(mods):
    Engine, capi, H = mods
    with pytest.raises(capi.OppError):
        Engine(46, 54, 20, 432)          # down-sampling is not INTER_AREA up-sampling
    with pytest.raises(capi.OppError):
        Engine(46, 54, gauss_kernel_size=16)
    eng = Engine(46, 54, max_batch=2)
    conf, paf = synth.render_batch(3, n_people=1)
    with pytest.raises(capi.OppError):
        eng.process(conf, paf)           # more frames than max_batch


def test_cpp_paf_processor_dropin(mods, tmp_path):
    """A C++ program using only the reference's API (create_paf_processor + operator()) linked against
    the library reproduces the reference build's golden humans."""
    import subprocess
    import conftest
    Engine, capi, H = mods
    root = conftest.ROOT
    exe = str(tmp_path / "dropin_main")
    libdir = os.path.join(root, "openpose_plus_b200")
    cmd = ["g++", "-std=c++14", "-O1", "-I", os.path.join(root, "include"), os.path.join(root, "tests", "cpp", "dropin_main.cpp"), "-o", exe,
           "-L", libdir, "-l:libopp_b200.so", "-Wl,-rpath," + libdir, "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
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


def test_device_outputs_and_channels_last_device_input(mods):
    """Results can stay on the device (out_mem = DEVICE) and channels-last device tensors are accepted."""
    import torch
    Engine, capi, H = mods
    conf, paf = synth.render_batch(5, n_people=4, seed0=900)
    eng = Engine(46, 54, max_batch=5)
    ref_h, ref_c, ref_f = eng.process(conf, paf)
    d_h = torch.zeros((5, eng.max_humans * 292), dtype=torch.uint8, device="cuda")
    d_c = torch.full((5,), -1, dtype=torch.int32, device="cuda")
    d_f = torch.full((5,), -1, dtype=torch.int32, device="cuda")
    dc = torch.from_numpy(np.ascontiguousarray(conf.transpose(0, 2, 3, 1))).cuda()
    dp = torch.from_numpy(np.ascontiguousarray(paf.transpose(0, 2, 3, 1))).cuda()
    eng.process(dc, dp, layout=capi.LAYOUT_HWC, out=(d_h, d_c, d_f))
    torch.cuda.synchronize()
    assert np.array_equal(d_c.cpu().numpy(), ref_c) and np.array_equal(d_f.cpu().numpy(), ref_f)
    got = d_h.cpu().numpy().view(capi.HUMAN_DT).reshape(5, eng.max_humans)
    for f in range(5):
        assert H.humans_equal(got[f, :ref_c[f]], ref_h[f, :ref_c[f]]) is None


def test_pageable_host_buffers_and_slot_reuse(mods):
    """Plain (pageable) numpy inputs/outputs take the staged path; many more batches than slots."""
    Engine, capi, H = mods
    conf, paf = synth.render_batch(6, n_people=3, seed0=950)
    eng = Engine(46, 54, max_batch=2, n_slots=2)
    ref = [eng.process(conf[i:i + 2], paf[i:i + 2]) for i in range(0, 6, 2)]
    tickets = []
    outs = []
    for rep in range(4):
        for i in range(0, 6, 2):
            if len(tickets) == 2:
                outs.append(eng.wait(tickets.pop(0)))
            tickets.append(eng.submit(conf[i:i + 2], paf[i:i + 2]))
    while tickets:
        outs.append(eng.wait(tickets.pop(0)))
    assert len(outs) == 12
    for k, (h, c, f) in enumerate(outs):
        rh, rc, rf = ref[k % 3]
        assert np.array_equal(c, rc) and np.array_equal(f, rf)
        for j in range(2):
            assert H.humans_equal(h[j, :c[j]], rh[j, :rc[j]]) is None
    with pytest.raises(capi.OppError):   # third submit without a wait: no free slot
        t1, t2 = eng.submit(conf[:2], paf[:2]), eng.submit(conf[:2], paf[:2])
        try:
            eng.submit(conf[:2], paf[:2])
        finally:
            eng.wait(t1), eng.wait(t2)


def test_two_devices_in_one_process(mods):
    """Handles on different GPUs of one process (kernel attributes are per device)."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    Engine, capi, H = mods
    conf, paf = synth.render_batch(2, n_people=3, seed0=970)
    e0, e1 = Engine(46, 54, max_batch=2, device=0), Engine(46, 54, max_batch=2, device=1)
    r0, r1 = e0.process(conf, paf), e1.process(conf, paf)
    assert e0.device == 0 and e1.device == 1
    assert np.array_equal(r0[1], r1[1])
    for f in range(2):
        assert H.humans_equal(r0[0][f, :r0[1][f]], r1[0][f, :r1[1][f]]) is None


def test_calls_leave_the_callers_current_device_alone(mods):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    Engine, capi, H = mods
    torch.cuda.set_device(0)
    conf, paf = synth.render_batch(1, n_people=2, seed0=980)
    e1 = Engine(46, 54, max_batch=1, device=1)
    assert torch.cuda.current_device() == 0
    e1.process(conf, paf)
    assert torch.cuda.current_device() == 0
    e1.close()
    assert torch.cuda.current_device() == 0


def test_single_process_multi_gpu_stream(mods):
    """configs[4] in one process: the stream sharded over every visible GPU equals the single-GPU result."""
    import torch
    from openpose_plus_b200.sharding import process_stream, process_stream_multi
    Engine, capi, H = mods
    ndev = torch.cuda.device_count()
    if ndev < 2:
        pytest.skip("needs two GPUs")
    base_c, base_p = synth.render_batch(8, n_people=4, seed0=990)
    idx = np.arange(1000) % 8
    conf, paf = base_c[idx], base_p[idx]
    engines = [Engine(46, 54, max_batch=32, device=d) for d in range(ndev)]
    humans, counts, flags = process_stream_multi(engines, conf, paf)
    h1, c1, f1 = process_stream(engines[0], conf, paf)
    assert np.array_equal(counts, c1) and np.array_equal(flags, f1)
    for f in range(0, 1000, 37):
        assert H.humans_equal(humans[f, :counts[f]], h1[f, :c1[f]]) is None


def test_latency_path_pipelined_single_frames(mods, monkeypatch):
    """One-frame batches in pinned memory take the latency path (ingest kernel for the heat maps, programmatic
    dependent launch, PAF tiles fetched from pinned memory by the limb kernel, completion word).  Three of them in
    flight on three slots, pinned and pageable result buffers, and every switch of that path turned off in turn must all
    give the batch path's answer."""
    Engine, capi, H = mods
    conf, paf = synth.render_batch(12, n_people=6, seed0=1200)
    ref_eng = Engine(46, 54, max_batch=12)
    want_h, want_c, want_f = ref_eng.process(conf, paf)
    hc, hp = capi.pinned_empty(conf.shape, np.float32), capi.pinned_empty(paf.shape, np.float32)
    hc[...] = conf
    hp[...] = paf

    def run(pinned_out):
        eng = Engine(46, 54, max_batch=4, n_slots=3)
        if pinned_out:
            outs = [(capi.pinned_empty((1, eng.max_humans), capi.HUMAN_DT), capi.pinned_empty((1,), np.int32), capi.pinned_empty((1,), np.int32)) for _ in range(12)]
        else:
            outs = [(np.zeros((1, eng.max_humans), capi.HUMAN_DT), np.zeros(1, np.int32), np.zeros(1, np.int32)) for _ in range(12)]
        inflight = []
        for f in range(12):
            if len(inflight) == 3:
                eng.wait(inflight.pop(0))
            inflight.append(eng.submit(hc[f:f + 1], hp[f:f + 1], out=outs[f]))
        for t in inflight:
            eng.wait(t)
        for f in range(12):
            h, c, fl = outs[f]
            assert c[0] == want_c[f] and fl[0] == want_f[f], f
            assert H.humans_equal(h[0, :c[0]], want_h[f, :want_c[f]]) is None, f
        t = eng.submit(hc[3:5], hp[3:5])                          # two frames: still the latency path
        two = eng.wait(t)
        assert np.array_equal(two[1], want_c[3:5]) and H.humans_equal(two[0][1, :two[1][1]], want_h[4, :want_c[4]]) is None
        assert 0 < eng.last_batch_ms(t) < 5                       # read from the events lazily, after the completion word
        eng.close()

    run(True)
    run(False)
    # pinned buffers that are only 4-byte aligned: the scalar ingest and the cp.async form of the PAF tile fetch
    raw_c, raw_p = capi.pinned_empty((conf[0].size + 1,), np.float32), capi.pinned_empty((paf[0].size + 1,), np.float32)
    oc, op = raw_c[1:].reshape(conf[:1].shape), raw_p[1:].reshape(paf[:1].shape)
    assert oc.ctypes.data % 16 == 4
    eng = Engine(46, 54, max_batch=1)
    for f in (0, 5):
        oc[...] = conf[f]
        op[...] = paf[f]
        h, c, fl = eng.process(oc, op)
        assert c[0] == want_c[f] and H.humans_equal(h[0, :c[0]], want_h[f, :want_c[f]]) is None, f
    eng.close()
    # plain numpy (pageable) inputs: staged through the slot's pinned buffers, PAFs copied while the GPU already works
    eng = Engine(46, 54, max_batch=4)
    for f in (1, 7):
        h, c, fl = eng.process(conf[f:f + 3], paf[f:f + 3])
        for k in range(3):
            assert c[k] == want_c[f + k] and H.humans_equal(h[k, :c[k]], want_h[f + k, :want_c[f + k]]) is None, (f, k)
    eng.close()
    for switch in ("OPP_NO_PDL", "OPP_NO_PAF_EARLY", "OPP_NO_DONE_FLAG"):
        monkeypatch.setenv(switch, "1")
        run(True)
        monkeypatch.delenv(switch)

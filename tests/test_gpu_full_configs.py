"""-m gpu: BASELINE.json's configurations at their full sizes.  configs[1]-[3] compare EVERY frame of the batch with the
CPU oracle (final skeletons bit-exact; the stage-by-stage comparison is test_gpu_parity.py's job); configs[4], 4096
frames, is checked through properties that do not need 4096 oracle runs: the stream equals the tiling of its distinct
frames' results (position in the stream, batch boundaries and pipeline slots must not matter) and is reproducible."""
import numpy as np
import pytest

from openpose_plus_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    from openpose_plus_b200.engine import Engine
    from oracle.oracle import Oracle, FLAG_UB_PEAK_INDEX
    import helpers
    return Engine, Oracle, helpers, FLAG_UB_PEAK_INDEX


def _check_batch(mods, conf, paf, eng, orc, what):
    Engine, Oracle, H, UB = mods
    from openpose_plus_b200 import _capi as capi
    humans, counts, flags = eng.process(conf, paf)
    assert not (flags & capi.FLAG_OVERFLOW_MASK).any(), what
    for f in range(conf.shape[0]):
        o = orc.run(conf[f], paf[f], lazy=True)
        if o["flags"] & UB:
            continue
        assert counts[f] == o["n_humans"], "%s frame %d: %d humans != %d" % (what, f, counts[f], o["n_humans"])
        err = H.humans_equal(humans[f, :counts[f]], o["humans"])
        assert err is None, "%s frame %d: %s" % (what, f, err)
    return counts


def test_config1_batch64_368x432_every_frame(mods):
    Engine, Oracle = mods[0], mods[1]
    conf, paf = synth.render_batch(64, n_people=5, seed0=5000)
    counts = _check_batch(mods, conf, paf, Engine(46, 54, max_batch=64), Oracle(46, 54, 368, 432, 17), "configs[1]")
    assert counts.sum() > 200


def test_config2_batch32_736x864_every_frame(mods):
    Engine, Oracle = mods[0], mods[1]
    conf, paf = synth.render_batch(32, n_people=12, feat_h=92, feat_w=108, seed0=5100)
    _check_batch(mods, conf, paf, Engine(92, 108, max_batch=32), Oracle(92, 108, 736, 864, 17), "configs[2]")


def test_config3_crowded_batch64_every_frame(mods):
    Engine, Oracle = mods[0], mods[1]
    fr = [synth.render_frame(5200 + i, n_people=30 + i % 9, drop_limbs=(12,) if i % 4 == 3 else ()) for i in range(64)]
    conf, paf = np.stack([f[0] for f in fr]), np.stack([f[1] for f in fr])
    counts = _check_batch(mods, conf, paf, Engine(46, 54, max_batch=64, max_humans=256), Oracle(46, 54, 368, 432, 17), "configs[3]")
    assert counts.mean() >= 25


def test_config4_stream_4096_frames_properties(mods):
    Engine, Oracle, H = mods[0], mods[1], mods[2]
    from openpose_plus_b200.sharding import process_stream
    base_c, base_p = synth.render_batch(13, n_people=6, seed0=5300)           # 13 distinct frames, period coprime with the batch
    idx = np.arange(4096) % 13
    conf, paf = base_c[idx], base_p[idx]
    eng = Engine(46, 54, max_batch=64)
    h0, c0, f0 = eng.process(base_c, base_p)                                  # the distinct frames on their own
    orc = Oracle(46, 54, 368, 432, 17)
    for f in range(13):                                                       # ... anchored to the oracle
        o = orc.run(base_c[f], base_p[f], lazy=True)
        assert c0[f] == o["n_humans"] and H.humans_equal(h0[f, :c0[f]], o["humans"]) is None
    humans, counts, flags = process_stream(eng, conf, paf)                    # 64 batches of 64 through 3 slots
    assert np.array_equal(counts, c0[idx]) and np.array_equal(flags, f0[idx])
    for f in range(4096):
        n = counts[f]
        assert np.array_equal(humans[f, :n].view(np.uint8), h0[idx[f], :n].view(np.uint8)), f
    again = process_stream(eng, conf, paf, batch=37)                          # ragged batches, same answer
    assert np.array_equal(again[1], counts)
    assert all(np.array_equal(again[0][f, :counts[f]].view(np.uint8), humans[f, :counts[f]].view(np.uint8)) for f in range(0, 4096, 97))

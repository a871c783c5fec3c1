"""Shared comparison helpers: CUDA path (through the C-ABI) against the CPU oracle."""
import numpy as np

from oracle.oracle import Oracle, FLAG_UB_PEAK_INDEX
from openpose_plus_b200 import _capi as capi


def bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def humans_equal(gpu, orc, score_rtol=0.0):
    """Field-wise comparison of human_t arrays.  Integer-valued fields (has_value, x, y) and peak
    scores must be bit-exact; human.score is bit-exact when score_rtol == 0, else within rtol."""
    if len(gpu) != len(orc):
        return "count %d != %d" % (len(gpu), len(orc))
    if len(gpu) == 0:
        return None
    if not np.array_equal(gpu["parts"]["has_value"] != 0, orc["parts"]["has_value"] != 0):
        return "has_value differs"
    for f in ("x", "y", "score"):
        if not np.array_equal(bits(gpu["parts"][f]), bits(orc["parts"][f])):
            return "parts.%s differs" % f
    if score_rtol == 0.0:
        if not np.array_equal(bits(gpu["score"]), bits(orc["score"])):
            return "human score bits differ"
    elif not np.allclose(gpu["score"], orc["score"], rtol=score_rtol, atol=0):
        return "human score beyond rtol"
    return None


def check_frame(engine, ticket, f, humans, counts, flags, o, what=""):
    """Every stage of frame f against oracle result o (dict from Oracle.run)."""
    tag = "%s frame %d: " % (what, f)
    assert (flags[f] & capi.FLAG_OVERFLOW_MASK) == 0, tag + "capacity overflow flags=%d" % flags[f]
    pk = engine.debug_peaks(ticket, f)
    op = o["peaks"]
    assert len(pk) == len(op), tag + "peak count %d != %d" % (len(pk), len(op))
    for fld in ("part_id", "x", "y", "id"):
        assert np.array_equal(pk[fld], op[fld]), tag + "peak %s differs" % fld
    assert np.array_equal(bits(pk["score"]), bits(op["score"])), tag + "peak score bits differ"
    for p in range(19):
        cn = engine.debug_conns(ticket, f, p)
        oc = o["conns"][p]
        assert len(cn) == len(oc), tag + "limb %d: %d conns != %d" % (p, len(cn), len(oc))
        assert np.array_equal(cn["cid1"], oc["cid1"]) and np.array_equal(cn["cid2"], oc["cid2"]), tag + "limb %d assignment differs" % p
        assert np.array_equal(bits(cn["score"]), bits(oc["score"])), tag + "limb %d PAF score bits differ" % p
    if o["flags"] & FLAG_UB_PEAK_INDEX:
        return  # the reference reads all_peaks[] out of bounds here: undefined, excluded
    n = int(counts[f])
    assert n == o["n_humans"], tag + "%d humans != %d" % (n, o["n_humans"])
    err = humans_equal(humans[f, :n], o["humans"])
    assert err is None, tag + err
    for i in range(n):
        assert np.array_equal(engine.debug_parts(ticket, f, i), o["hrefs"]["parts"][i]), tag + "human %d part ids differ" % i
    c = engine.debug_counts(ticket, f)
    assert c[0] == o["n_incomplete"] and c[1] == o["n_merges"], tag + "assembly counters differ"


def run_and_check(engine, oracle, conf, paf, what="", **kw):
    t = engine.submit(conf, paf, **kw)
    humans, counts, flags = engine.wait(t)
    for f in range(conf.shape[0]):
        o = oracle.run(conf[f], paf[f])
        check_frame(engine, t, f, humans, counts, flags, o, what)
    return humans, counts, flags

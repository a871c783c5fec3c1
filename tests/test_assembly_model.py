"""CPU: an executable model of the limb kernel's assembly (csrc/opp_kernels.cu assemble_frame) checked against the oracle.

The CUDA assembly does not replay src/paf.cpp:177-262 connection by connection; it relies on three claims:
  (1) the 17 tree limbs can be evaluated as a forest: creators found from the parent limb / earlier limbs with the same
      first part, numbered in (limb, connection) order, parts settled by three relaxation rounds, scores summed limb
      by limb per human;
  (2) a prefix of each virtual limb (until the first connection that touches two humans, or one human holding another
      peak in that part) can be applied in any order;
  (3) from there on the reference's sequential rule - stored-id indexing, `> 0` membership, erase - finishes the limb.
This model implements exactly that decomposition in numpy/Python from the oracle's connections and must reproduce the
oracle's partial humans (ids, part ids, scores bit for bit, part counts) on typical, crowded, merge-heavy frames."""
import numpy as np
import pytest

from oracle.oracle import Oracle, FLAG_UB_STALE_INDEX, FLAG_UB_PEAK_INDEX
from openpose_plus_b200 import synth

PA = [1, 1, 2, 3, 5, 6, 1, 8, 9, 1, 11, 12, 1, 0, 14, 0, 15, 2, 5]
PB = [2, 5, 3, 4, 6, 7, 8, 9, 10, 11, 12, 13, 0, 14, 16, 15, 17, 16, 17]
f32 = np.float32


def forest(conns, ps, pofs):
    flat = [(l, int(c["cid1"]), int(c["cid2"]), f32(c["score"])) for l in range(17) for c in conns[l]]
    c1 = {(l, a): t for t, (l, a, b, s) in enumerate(flat)}
    brought = {b for (l, a, b, s) in flat}
    creators = []
    for t, (l, a, b, s) in enumerate(flat):
        if PA[l] == 1:
            owned = any((lp, a) in c1 for lp in (0, 1, 6, 9) if lp < l)
        else:
            owned = a in brought or (l == 15 and (13, a) in c1)
        if not owned:
            creators.append(t)
    contrib = [f32(f32(ps(a) + ps(b)) + s) if t in set(creators) else f32(ps(b) + s) for t, (l, a, b, s) in enumerate(flat)]
    humans = []
    for q, t0 in enumerate(creators):
        l0, a, b, _ = flat[t0]
        parts = [-1] * 18
        parts[PA[l0]], parts[PB[l0]] = a, b
        humans.append(dict(id=q, parts=parts, t0=t0, l0=l0))
    for _ in range(3):  # relaxation rounds over (human, limb)
        for h in humans:
            for l in range(17):
                held = h["parts"][PA[l]]
                if held < 0 or h["parts"][PB[l]] != -1:
                    continue
                t = c1.get((l, held))
                if t is not None:
                    h["parts"][PB[l]] = flat[t][2]
    for h in humans:
        sc, n = contrib[h["t0"]], 2
        for l in range(17):
            held = h["parts"][PA[l]]
            if l == h["l0"] or held < 0:
                continue
            t = c1.get((l, held))
            if t is not None:
                sc, n = f32(sc + contrib[t]), n + 1
        h["score"], h["n"] = sc, n
    return humans


def clone(h):
    return dict(h, parts=list(h["parts"]))


def virtual_limb(mem, conns_l, l, ps, state):
    """mem = the vector's storage (slots beyond state['n'] keep their old bytes, as after std::vector::erase);
    parallel prefix while nothing was erased, then the reference's sequential rule (stored ids, erase)."""
    p1, p2 = PA[l], PB[l]

    def extend(h, c):
        h["parts"][p2] = int(c["cid2"])
        h["n"] += 1
        h["score"] = f32(h["score"] + f32(ps(int(c["cid2"])) + f32(c["score"])))

    k0 = 0
    if state["merges"] == 0:
        live = mem[:state["n"]]
        hits = [[q for q, h in enumerate(live) if h["parts"][p1] == c["cid1"] or h["parts"][p2] == c["cid2"]] for c in conns_l]
        k0 = len(conns_l)
        for k, (c, hs) in enumerate(zip(conns_l, hits)):
            cur = live[hs[0]]["parts"][p2] if len(hs) == 1 else None
            if not (len(hs) == 0 or (len(hs) == 1 and cur in (c["cid2"], -1))):
                k0 = k
                break
        for k in reversed(range(k0)):  # the safe ones, in any order (here: backwards, to make the point)
            c, hs = conns_l[k], hits[k]
            if len(hs) == 1 and live[hs[0]]["parts"][p2] == -1:
                extend(live[hs[0]], c)
    for c in conns_l[k0:]:  # src/paf.cpp:192-248 for a virtual pair
        # src/paf.cpp:198 collects the STORED ids (stale after an erase); the Python path's pafprocess the positions
        ids = [(q if state["true_index"] else h["id"]) for q, h in enumerate(mem[:state["n"]]) if h["parts"][p1] == c["cid1"] or h["parts"][p2] == c["cid2"]]
        if len(ids) == 1:
            if ids[0] >= state["hist_max"]:
                state["ub"] = True
                continue
            if mem[ids[0]]["parts"][p2] != c["cid2"]:
                extend(mem[ids[0]], c)
        elif len(ids) >= 2:
            if max(ids[0], ids[1]) >= state["hist_max"]:
                state["ub"] = True
                continue
            h1, h2 = mem[ids[0]], mem[ids[1]]
            if any(h1["parts"][i] > 0 and h2["parts"][i] > 0 for i in range(18)):
                extend(h1, c)
            else:
                for i in range(18):
                    h1["parts"][i] += h2["parts"][i] + 1
                h1["n"] += h2["n"]
                h1["score"] = f32(f32(h1["score"] + h2["score"]) + f32(c["score"]))
                e = ids[1]
                if e < state["n"]:
                    for q in range(e, state["n"] - 1):
                        mem[q] = clone(mem[q + 1])
                else:
                    state["ub"] = True  # erase at or past end(): undefined in the reference; the CUDA tests cover what it does
                state["n"] -= 1
                state["merges"] += 1


def model(o, true_index=False):
    peaks = o["peaks"]
    ps = lambda i: f32(peaks["score"][i])  # noqa: E731
    pofs = np.concatenate([[0], np.cumsum([(peaks["part_id"] == k).sum() for k in range(18)])])
    mem = forest(o["conns"], ps, pofs)
    state = dict(merges=0, n=len(mem), hist_max=len(mem), ub=False, true_index=true_index)
    for l in (17, 18):
        virtual_limb(mem, list(o["conns"][l]), l, ps, state)
    return mem[:state["n"]], state


def frames():
    for seed in range(20):
        yield synth.render_frame(700 + seed, n_people=3 + seed % 7)
    for seed in range(14):
        yield synth.render_frame(800 + seed, n_people=24 + seed, drop_limbs=(12,) if seed % 2 else (), noise=1e-3 if seed % 3 == 0 else 0.0)
    for seed in range(6):
        yield synth.render_frame(900 + seed, n_people=12, drop_limbs=(int(seed * 3 % 17),))


def test_forest_and_virtual_prefix_model_reproduces_the_oracle():
    orc = Oracle(46, 54, 368, 432, 17)
    checked = merged = 0
    for conf, paf in frames():
        o = orc.run(conf, paf, lazy=True)
        humans, st = model(o)
        if st["ub"] or (o["flags"] & (FLAG_UB_STALE_INDEX | FLAG_UB_PEAK_INDEX | 4)):
            continue  # frames where the reference indexes / erases out of range are the CUDA tests' business
        assert len(humans) == o["n_incomplete"] and st["merges"] == o["n_merges"]
        keep = [h for h in humans if not (h["n"] < 4 or f32(h["score"] / f32(h["n"])) < f32(0.4))]
        assert len(keep) == o["n_humans"]
        for h, r in zip(keep, o["hrefs"]):
            assert h["parts"] == r["parts"].tolist() and h["n"] == r["n_parts"]
            assert np.float32(h["score"]).view(np.uint32) == np.float32(r["score"]).view(np.uint32)
        checked += 1
        merged += st["merges"]
    assert checked >= 25 and merged >= 20


def test_position_indexed_model_reproduces_oracle_variant_1():
    """The same decomposition with humans indexed by position (OPP_VARIANT_PYTHON, the pafprocess rule): an independent
    statement of the grouping half of oracle variant 1.  Positions are never stale, so no frame is skipped."""
    orc = Oracle(46, 54, 368, 432, 25, variant=1)
    merged = 0
    for n_frames, (conf, paf) in enumerate(frames()):
        o = orc.run(conf, paf, lazy=True)
        humans, st = model(o, true_index=True)
        assert not st["ub"] and o["flags"] == 0
        assert len(humans) == o["n_incomplete"] and st["merges"] == o["n_merges"]
        keep = [h for h in humans if not (h["n"] < 4 or f32(h["score"] / f32(h["n"])) < f32(0.4))]
        assert len(keep) == o["n_humans"]
        for h, r in zip(keep, o["hrefs"]):
            assert h["parts"] == r["parts"].tolist() and h["n"] == r["n_parts"]
            assert np.float32(h["score"]).view(np.uint32) == np.float32(r["score"]).view(np.uint32)
        merged += st["merges"]
    assert n_frames >= 39 and merged >= 50

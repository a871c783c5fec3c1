#!/usr/bin/env python
"""Mints tests/golden/pyvariant_*.npz: regression vectors for OPP_VARIANT_PYTHON (the semantics of the reference's Python
graph, openpose_plus/inference/post_process.py:13-37,82-106).  PARITY UNPINNED: neither TensorFlow nor the external
pafprocess module is available, so these vectors come from this repo's own restatement (oracle variant 1) after its
smoothing was checked against a float64 2-D convolution with the reference's _gauss_kernel; they guard against
regressions, they are not reference outputs."""
import os
import sys

import numpy as np
import scipy.signal
import scipy.stats

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

from oracle.oracle import Oracle  # noqa: E402
from openpose_plus_b200 import synth  # noqa: E402


def gauss_kernel(ksize, nsig):  # post_process.py:13-17, restated
    interval = (2 * nsig + 1.) / ksize
    x = np.linspace(-nsig - interval / 2., nsig + interval / 2., ksize + 1)
    y = np.diff(scipy.stats.norm.cdf(x))
    k = np.sqrt(np.outer(y, y))
    return k / k.sum()


def case(name, conf, paf, out_h, out_w, ksize):
    h, w = conf.shape[1:]
    o = Oracle(h, w, out_h, out_w, ksize, variant=1).run(conf, paf, maps=True)
    want = scipy.signal.convolve2d(o["conf_up"][3].astype(np.float64), gauss_kernel(ksize, 3.0), mode="same", boundary="fill")
    assert np.abs(o["smoothed"][3] - want).max() < 1e-6
    d = dict(conf=conf, paf=paf, geom=np.array([h, w, out_h, out_w, ksize], np.int32), humans=o["humans"], peaks=o["peaks"], hrefs=o["hrefs"],
             counts=np.array([o["n_incomplete"], o["n_merges"], o["flags"]], np.int32))
    for p in range(19):
        d["conns_%02d" % p] = o["conns"][p]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "peaks", len(o["peaks"]), "humans", o["n_humans"], "merges", o["n_merges"], "flags", o["flags"])


if __name__ == "__main__":
    case("pyvariant_6p_368x432_k25", *synth.render_frame(15, 6), 368, 432, 25)
    case("pyvariant_33p_merge_368x432_k25", *synth.render_frame(203, 33, drop_limbs=(12,)), 368, 432, 25)

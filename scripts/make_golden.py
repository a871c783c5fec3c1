#!/usr/bin/env python
"""Mints the golden vectors under tests/golden/ (run HERE, where /root/reference and cv2 exist).

The reference holds no fixtures for this path (CMakeLists.txt:24-25 "# TODO: add tests"), so the
vectors are generated from outputs of the reference itself:
  * humans_ref  -- oracle/_ref/libopp_ref.so = the reference's own unmodified src/paf.cpp (strict
                   build) run on the stored inputs;
  * cv_*        -- cv2 4.13 with IPP off and setUseOptimized(False), for the two OpenCV calls the
                   reference makes (cv::resize INTER_AREA, cv::GaussianBlur sigma=3) and
                   cv::getGaussianKernel;
  * peaks/conns/hrefs -- the instrumented C restatement (oracle/liborc.so), stored only after its
                   final humans were checked identical to humans_ref.
Inputs are stored with the outputs so the vectors do not depend on numpy/libm versions.
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, "tests", "golden")

import cv2  # noqa: E402

cv2.ipp.setUseIPP(False)
cv2.setUseOptimized(False)
cv2.setNumThreads(1)

from oracle.oracle import Oracle, Reference  # noqa: E402
from openpose_plus_b200 import synth  # noqa: E402


def same_humans(a, b):
    if len(a) != len(b):
        return False
    ok = np.array_equal(a["score"].view(np.uint32), b["score"].view(np.uint32))
    ok &= np.array_equal(a["parts"]["has_value"] != 0, b["parts"]["has_value"] != 0)
    for f in ("x", "y", "score"):
        ok &= np.array_equal(np.ascontiguousarray(a["parts"][f]).view(np.uint32), np.ascontiguousarray(b["parts"][f]).view(np.uint32))
    return bool(ok)


def frame_case(name, conf, paf, out_h, out_w, ksize):
    h, w = conf.shape[1:]
    orc, ref = Oracle(h, w, out_h, out_w, ksize), Reference(h, w, out_h, out_w, ksize)
    o, r = orc.run(conf, paf), ref.run(conf, paf, cap=8192)
    assert same_humans(o["humans"], r), name + ": restatement disagrees with the reference build"
    d = dict(conf=conf, paf=paf, geom=np.array([h, w, out_h, out_w, ksize], np.int32), humans_ref=r, peaks=o["peaks"], hrefs=o["hrefs"],
             counts=np.array([o["n_incomplete"], o["n_merges"], o["flags"]], np.int32))
    for p in range(19):
        d["conns_%02d" % p] = o["conns"][p]
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **d)
    print(name, "peaks", len(o["peaks"]), "humans", len(r), "merges", o["n_merges"], "flags", o["flags"], "ties", sum(o["ties"]))


def main():
    os.makedirs(OUT, exist_ok=True)
    frame_case("frame_5p_368x432_k17", *synth.render_frame(11, 5), 368, 432, 17)
    frame_case("frame_8p_368x432_k9", *synth.render_frame(12, 8), 368, 432, 9)
    frame_case("frame_35p_368x432_k13", *synth.render_frame(2, 35), 368, 432, 13)
    frame_case("frame_40p_merge_368x432_k17", *synth.render_frame(5, 40, drop_limbs=(12,)), 368, 432, 17)
    frame_case("frame_12p_736x864_k17", *synth.render_frame(13, 12, 92, 108), 736, 864, 17)
    frame_case("frame_6p_300x400_k17", *synth.render_frame(14, 6), 300, 400, 17)
    frame_case("frame_6p_368x432_k25", *synth.render_frame(15, 6), 368, 432, 25)
    frame_case("frame_noise_96x112_k17", *synth.noise_frame(7, 12, 14), 96, 112, 17)
    frame_case("frame_empty_368x432_k17", np.zeros((19, 46, 54), np.float32), np.zeros((38, 46, 54), np.float32), 368, 432, 17)

    # OpenCV pieces
    rng = np.random.default_rng(123)
    d = {}
    for k in range(1, 64, 2):
        d["taps_%d" % k] = cv2.getGaussianKernel(k, 3.0, cv2.CV_32F).ravel()
    src = (rng.random((23, 27), dtype=np.float32) * 2 - 1).astype(np.float32)
    d["resize_src"] = src
    for (H, W) in [(184, 216), (100, 97), (23, 27), (23, 54), (47, 55), (69, 81)]:
        d["resize_%dx%d" % (H, W)] = cv2.resize(src, (W, H), interpolation=cv2.INTER_AREA)
    blur_src = np.repeat(np.repeat(rng.random((12, 14), dtype=np.float32), 8, 0), 8, 1)
    d["blur_src"] = blur_src
    d["blur_src2"] = rng.random((40, 33), dtype=np.float32)
    for k in (1, 3, 5, 7, 9, 13, 17, 25, 31):
        d["blur_%d" % k] = cv2.GaussianBlur(blur_src, (k, k), 3.0)
        d["blur2_%d" % k] = cv2.GaussianBlur(d["blur_src2"], (k, k), 3.0)
    d["dilate"] = cv2.dilate(d["blur_src2"], np.ones((3, 3), np.uint8))
    np.savez_compressed(os.path.join(OUT, "opencv_pieces.npz"), **d)
    print("opencv pieces", len(d))


if __name__ == "__main__":
    main()

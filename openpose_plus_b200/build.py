"""Builds the in-tree CUDA library (sm_100a only) with nvcc; no torch involved.

    python -m openpose_plus_b200.build [--force]

Output: openpose_plus_b200/libopp_b200.so (git-ignored; it travels to the GPU box with the snapshot).
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libopp_b200.so")
SOURCES = ["opp_kernels.cu", "opp_capi.cu", "paf_processor.cpp", "vis.cpp"]
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
# -fmad=false: every float operation that decides an output must round exactly like the
# reference's strict-IEEE scalar code; nothing on this path wants a fused multiply-add.
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-Wall", "-shared", "-I", os.path.join(ROOT, "include")]


def stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "opp_b200.h"), __file__]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False):
    if not force and not stale():
        return OUT
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT]
    subprocess.run(cmd, check=True)
    check_sass()
    return OUT


OUT_DEBUG = os.path.join(HERE, "libopp_b200_dbg.so")


def build_debug(force=False):
    """The same library with every shared-memory array access bounds-checked (-DOPP_DEBUG_BOUNDS, csrc/opp_kernels.cu Span):
    the stand-in for compute-sanitizer, which the GPU pool does not offer.  Load it with OPP_B200_LIB=<path>; the fuzzers in
    scripts/ report what opp_debug_bounds_report recorded."""
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)] + [os.path.join(ROOT, "include", "opp_b200.h"), __file__]
    if not force and os.path.exists(OUT_DEBUG) and all(os.path.getmtime(d) <= os.path.getmtime(OUT_DEBUG) for d in deps):
        return OUT_DEBUG
    flags = [f for f in FLAGS if f != "-O3"] + ["-O2", "-DOPP_DEBUG_BOUNDS"]
    subprocess.run([NVCC] + flags + [os.path.join(CSRC, s) for s in SOURCES] + ["-o", OUT_DEBUG], check=True)
    return OUT_DEBUG


CUOBJDUMP = os.path.join(os.path.dirname(NVCC), "cuobjdump")


def fused_multiply_adds(lib=None):
    """{kernel: count} of fused multiply-adds (FFMA, FFMA2, HFMA2 excluded) in the SASS of the resize / peak kernels.
    Bit-exactness with the reference's scalar filter rests on every product being rounded before it is added: the
    library is built with -fmad=false and the packed adds (add.rn.f32x2) are fed by scalar multiplications only, but
    ptxas is known to contract a packed multiply feeding a packed add into FFMA2 regardless, so the SASS is checked.
    (The limb kernel's IEEE divisions and square roots expand to FFMA sequences: those are the correctly rounded forms.)"""
    out = subprocess.run([CUOBJDUMP, "-sass", lib or OUT], check=True, capture_output=True, text=True).stdout
    counts, fn = {}, None
    for line in out.splitlines():
        if "Function :" in line:
            fn = line.split("Function :")[1].strip()
        elif fn and ("FFMA" in line or "DFMA" in line) and any(k in fn for k in ("k0_", "k1_", "k2_")):
            counts[fn] = counts.get(fn, 0) + 1
    return counts


def check_sass():
    bad = fused_multiply_adds()
    if bad:
        raise RuntimeError("fused multiply-adds in kernels that must round every product: %r" % bad)


REF_INCLUDE = "/root/reference/include"
DROPIN_REF = os.path.join(ROOT, "tests", "cpp", "_build", "dropin_refhdr")


def build_dropin_against_reference_headers(force=False):
    """tests/cpp/dropin_main.cpp - a caller written against the reference's public API only - compiled against the
    reference's OWN, unmodified headers (not this repo's re-written ones) and linked with libopp_b200.so: pins the ABI
    claim (vtable order of paf_processor, human_t layout, factory signature).  The reference tree exists only in the
    build container; the binary (git-ignored) travels to the GPU box, where the -m gpu tests run it.  Returns the path,
    or None when the reference headers are absent and no prebuilt binary exists."""
    src = os.path.join(ROOT, "tests", "cpp", "dropin_main.cpp")
    if not os.path.exists(os.path.join(REF_INCLUDE, "openpose-plus.hpp")):
        return DROPIN_REF if os.path.exists(DROPIN_REF) else None
    if not force and os.path.exists(DROPIN_REF) and os.path.getmtime(DROPIN_REF) >= max(os.path.getmtime(src), os.path.getmtime(OUT)):
        return DROPIN_REF
    os.makedirs(os.path.dirname(DROPIN_REF), exist_ok=True)
    # `-include string`: the reference's openpose-plus.hpp uses std::string without including it
    cmd = ["g++", "-std=c++14", "-O1", "-include", "string", "-I", REF_INCLUDE, src, "-o", DROPIN_REF, "-L", HERE, "-l:libopp_b200.so",
           "-Wl,-rpath,$ORIGIN/../../../openpose_plus_b200", "-L/usr/local/cuda/lib64", "-Wl,-rpath,/usr/local/cuda/lib64"]
    subprocess.run(cmd, check=True)
    return DROPIN_REF


if __name__ == "__main__":
    if "--debug" in sys.argv:
        print(build_debug(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))

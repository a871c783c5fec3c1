"""openpose-plus post-processing (feature maps in, grouped COCO-18 skeletons out), B200-native.

The compute path is the CUDA library built by `python -m openpose_plus_b200.build`
(openpose_plus_b200/libopp_b200.so, C-ABI in include/opp_b200.h).  Importing the package is cheap;
the library is loaded on first use and its absence is an error (there is no CPU path).
"""
from . import synth  # noqa: F401

__all__ = ["PostProcessor", "Human", "BodyPart", "Engine"]


def __getattr__(name):
    if name in ("PostProcessor", "Human", "BodyPart"):
        from . import post_process
        return getattr(post_process, name)
    if name == "Engine":
        from .engine import Engine
        return Engine
    raise AttributeError(name)

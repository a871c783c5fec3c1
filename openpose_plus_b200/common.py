"""Result-side helpers matching the reference's openpose_plus/inference/common.py (CocoPart :15-34,
CocoPairs :36-38, tranform_keypoints2d :86-96)."""
from enum import Enum

import numpy as np


class CocoPart(Enum):
    Nose = 0
    Neck = 1
    RShoulder = 2
    RElbow = 3
    RWrist = 4
    LShoulder = 5
    LElbow = 6
    LWrist = 7
    RHip = 8
    RKnee = 9
    RAnkle = 10
    LHip = 11
    LKnee = 12
    LAnkle = 13
    REye = 14
    LEye = 15
    REar = 16
    LEar = 17
    Background = 18


CocoPairs = [(1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10), (1, 11), (11, 12), (12, 13),
             (1, 0), (0, 14), (14, 16), (0, 15), (15, 17), (2, 16), (5, 17)]
CocoPairsRender = CocoPairs[:-2]


def keypoints_array(human, image_w, image_h):
    """[18, 3] array of (x px, y px, score), zeros where a part is missing."""
    out = np.zeros((18, 3), np.float32)
    for idx, bp in human.body_parts.items():
        out[idx] = (bp.x * image_w, bp.y * image_h, bp.score)
    return out

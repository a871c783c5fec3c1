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


def tranform_keypoints2d(body, width, height, kp_score_thresh=0.25):
    """(sic) The reference's helper of that name (openpose_plus/inference/common.py:86-96): `body` is a
    Human.body_parts dict {part_idx: BodyPart}; returns (coords2d [18,2] float64 in pixels, coords2d_conf [18] float64,
    coords2d_vis [18] bool = score > kp_score_thresh)."""
    coords2d = np.zeros((18, 2))
    coords2d_conf = np.zeros((18))
    coords2d_vis = coords2d_conf > 0.0
    for i, part in body.items():
        coords2d_conf[i] = part.score
        coords2d[i, 0] = part.x * width
        coords2d[i, 1] = part.y * height
        coords2d_vis[i] = coords2d_conf[i] > kp_score_thresh
    return coords2d, coords2d_conf, coords2d_vis


def keypoints_array(human, image_w, image_h):
    """[18, 3] array of (x px, y px, score), zeros where a part is missing."""
    out = np.zeros((18, 3), np.float32)
    for idx, bp in human.body_parts.items():
        out[idx] = (bp.x * image_w, bp.y * image_h, bp.score)
    return out


CocoColors = [(255, 0, 0), (255, 85, 0), (255, 170, 0), (255, 255, 0), (170, 255, 0), (85, 255, 0), (0, 255, 0), (0, 255, 85),
              (0, 255, 170), (0, 255, 255), (0, 170, 255), (0, 85, 255), (0, 0, 255), (85, 0, 255), (170, 0, 255), (255, 0, 255),
              (255, 0, 170), (255, 0, 85)]


def _disc(img, cx, cy, r, color):
    h, w = img.shape[:2]
    y0, y1, x0, x1 = max(cy - r, 0), min(cy + r + 1, h), max(cx - r, 0), min(cx + r + 1, w)
    if y0 >= y1 or x0 >= x1:
        return
    yy, xx = np.ogrid[y0:y1, x0:x1]
    img[y0:y1, x0:x1][(yy - cy) ** 2 + (xx - cx) ** 2 <= r * r] = color


def _line(img, p0, p1, color, thickness):
    n = int(max(abs(p1[0] - p0[0]), abs(p1[1] - p0[1]))) + 1
    xs = np.linspace(p0[0], p1[0], n).round().astype(int)
    ys = np.linspace(p0[1], p1[1], n).round().astype(int)
    r = max(thickness // 2, 0)
    for x, y in zip(xs, ys):
        _disc(img, int(x), int(y), r, color)


def draw_humans(npimg, humans, imgcopy=False):
    """Skeleton overlay like the reference's draw_humans (openpose_plus/inference/common.py:144-165) and
    draw_human (examples/vis.cpp:56-81): a dot per detected part, a line per rendered limb.  numpy only."""
    if imgcopy:
        npimg = np.copy(npimg)
    image_h, image_w = npimg.shape[:2]
    for human in humans:
        centers = {}
        for i in range(18):
            if i not in human.body_parts:
                continue
            bp = human.body_parts[i]
            centers[i] = (int(bp.x * image_w + 0.5), int(bp.y * image_h + 0.5))
            _disc(npimg, centers[i][0], centers[i][1], 3, CocoColors[i])
        for pair_order, (a, b) in enumerate(CocoPairsRender):
            if a in centers and b in centers:
                _line(npimg, centers[a], centers[b], CocoColors[pair_order], 3)
    return npimg

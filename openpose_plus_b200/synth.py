"""Synthetic part-confidence / part-affinity maps of the shape the openpose-plus CNN emits.

The CNN backbone is out of scope (BASELINE.json), so benches and tests feed maps rendered the way
the reference renders its own training targets (openpose_plus/utils.py:458-617): a Gaussian blob
per joint (sigma 8 px, cut at d^2/(2 sigma^2) > 4.6052, max over people), background = 1 - max,
unit limb vectors inside a +-8 px band around each limb segment (averaged where people overlap),
all drawn at image resolution and then area-averaged down by the network stride.

Channel conventions follow include/openpose-plus/coco.h:11-53: heat channel = COCO part index
(18 = background); limb ``pair_id`` (COCOPAIRS) writes PAF channels COCOPAIRS_NET[pair_id].
"""
import numpy as np

N_PARTS, N_PAIRS = 18, 19
COCOPAIRS = [(1, 2), (1, 5), (2, 3), (3, 4), (5, 6), (6, 7), (1, 8), (8, 9), (9, 10), (1, 11),
             (11, 12), (12, 13), (1, 0), (0, 14), (14, 16), (0, 15), (15, 17), (2, 16), (5, 17)]
COCOPAIRS_NET = [(12, 13), (20, 21), (14, 15), (16, 17), (22, 23), (24, 25), (0, 1), (2, 3), (4, 5), (6, 7),
                 (8, 9), (10, 11), (28, 29), (30, 31), (34, 35), (32, 33), (36, 37), (18, 19), (26, 27)]

# A standing figure in a unit box (x right, y down), COCO-18 order.
_TEMPLATE = np.array([
    (0.50, 0.08), (0.50, 0.20), (0.38, 0.20), (0.33, 0.36), (0.30, 0.50), (0.62, 0.20),
    (0.67, 0.36), (0.70, 0.50), (0.43, 0.52), (0.42, 0.74), (0.41, 0.95), (0.57, 0.52),
    (0.58, 0.74), (0.59, 0.95), (0.47, 0.06), (0.53, 0.06), (0.44, 0.08), (0.56, 0.08)], dtype=np.float64)

_SIGMA = 8.0
_CUT = 4.6052
_BAND = 8.0


def _skeletons(rng, n_people, H, W):
    out = []
    for _ in range(n_people):
        size = rng.uniform(0.25, 0.5) * H
        ox = rng.uniform(-0.1 * size, W - 0.9 * size)
        oy = rng.uniform(-0.05 * size, H - 0.95 * size)
        flip = rng.random() < 0.5
        t = _TEMPLATE.copy()
        if flip:
            t[:, 0] = 1.0 - t[:, 0]
        pts = np.stack([ox + t[:, 0] * size * 0.6 + 0.2 * size, oy + t[:, 1] * size], axis=1)
        pts += rng.normal(0.0, 2.0, pts.shape)
        out.append(pts)
    return out


def _draw_heat(heat, part, cx, cy):
    H, W = heat.shape[1:]
    d = np.sqrt(_CUT * 2) * _SIGMA
    x0, y0 = int(max(0, cx - d + 0.5)), int(max(0, cy - d + 0.5))
    x1, y1 = int(min(W - 1, cx + d + 0.5)), int(min(H - 1, cy + d + 0.5))
    if x1 < x0 or y1 < y0:
        return
    ys = (np.arange(y0, y1 + 1, dtype=np.float64) - cy) ** 2
    xs = (np.arange(x0, x1 + 1, dtype=np.float64) - cx) ** 2
    e = (ys[:, None] + xs[None, :]) / (2.0 * _SIGMA * _SIGMA)
    v = np.exp(-e)
    v[e > _CUT] = 0
    patch = heat[part, y0:y1 + 1, x0:x1 + 1]
    np.maximum(patch, v.astype(np.float32), out=patch)


def _draw_limb(vec, cnt, pair_id, a, b):
    H, W = cnt.shape[1:]
    vx, vy = b[0] - a[0], b[1] - a[1]
    length = float(np.hypot(vx, vy))
    if length == 0:
        return
    nx, ny = vx / length, vy / length
    x0, y0 = max(0, int(min(a[0], b[0]) - _BAND)), max(0, int(min(a[1], b[1]) - _BAND))
    x1, y1 = min(W, int(max(a[0], b[0]) + _BAND)), min(H, int(max(a[1], b[1]) + _BAND))
    if x1 <= x0 or y1 <= y0:
        return
    xs = np.arange(x0, x1, dtype=np.float64) - a[0]
    ys = np.arange(y0, y1, dtype=np.float64) - a[1]
    dist = np.abs(xs[None, :] * ny - ys[:, None] * nx)
    m = dist <= _BAND
    cx, cy = COCOPAIRS_NET[pair_id]
    vec[cx, y0:y1, x0:x1][m] += np.float32(nx)
    vec[cy, y0:y1, x0:x1][m] += np.float32(ny)
    cnt[pair_id, y0:y1, x0:x1][m] += 1


def _area_down(a, s):
    C, H, W = a.shape
    return a.reshape(C, H // s, s, W // s, s).mean(axis=(2, 4), dtype=np.float64).astype(np.float32)


def render_frame(seed, n_people=5, feat_h=46, feat_w=54, stride=8, noise=0.0, drop_limbs=()):
    """One frame: (conf [19,h,w] float32 in [0,1], paf [38,h,w] float32 in [-1,1])."""
    rng = np.random.default_rng(seed)
    H, W = feat_h * stride, feat_w * stride
    heat = np.zeros((N_PARTS + 1, H, W), np.float32)
    vec = np.zeros((2 * N_PAIRS, H, W), np.float32)
    cnt = np.zeros((N_PAIRS, H, W), np.int32)
    for pts in _skeletons(rng, n_people, H, W):
        for part in range(N_PARTS):
            _draw_heat(heat, part, pts[part, 0], pts[part, 1])
        for pair_id, (pa, pb) in enumerate(COCOPAIRS):
            if pair_id in drop_limbs:
                continue
            _draw_limb(vec, cnt, pair_id, pts[pa], pts[pb])
    heat[N_PARTS] = np.clip(1.0 - heat[:N_PARTS].max(axis=0), 0.0, 1.0)
    for pair_id, (cx, cy) in enumerate(COCOPAIRS_NET):
        c = np.maximum(cnt[pair_id], 1).astype(np.float32)
        vec[cx] /= c
        vec[cy] /= c
    conf, paf = _area_down(heat, stride), _area_down(vec, stride)
    if noise > 0:
        conf = np.clip(conf + rng.uniform(-noise, noise, conf.shape).astype(np.float32), 0, 1).astype(np.float32)
        paf = np.clip(paf + rng.uniform(-noise, noise, paf.shape).astype(np.float32), -1, 1).astype(np.float32)
    return np.ascontiguousarray(conf), np.ascontiguousarray(paf)


def noise_frame(seed, feat_h=46, feat_w=54):
    """Dense-noise stress frame: uniform [0,1) heat maps and uniform [-1,1) PAFs (plateau / tie stress)."""
    rng = np.random.default_rng(seed)
    conf = rng.random((N_PARTS + 1, feat_h, feat_w), dtype=np.float32)
    paf = (rng.random((2 * N_PAIRS, feat_h, feat_w), dtype=np.float32) * 2 - 1).astype(np.float32)
    return conf, paf


def render_batch(n_frames, n_people=5, feat_h=46, feat_w=54, stride=8, seed0=0, noise=0.0, pool=None):
    """(conf [n,19,h,w], paf [n,38,h,w]); frame i uses seed seed0+i.  With ``pool`` only that many
    distinct frames are rendered and then tiled (bench inputs: rendering is not what is measured)."""
    k = n_frames if pool is None else min(pool, n_frames)
    fr = [render_frame(seed0 + i, n_people, feat_h, feat_w, stride, noise) for i in range(k)]
    conf = np.stack([f[0] for f in fr])
    paf = np.stack([f[1] for f in fr])
    if k < n_frames:
        reps = -(-n_frames // k)
        conf = np.tile(conf, (reps, 1, 1, 1))[:n_frames]
        paf = np.tile(paf, (reps, 1, 1, 1))[:n_frames]
    return np.ascontiguousarray(conf), np.ascontiguousarray(paf)

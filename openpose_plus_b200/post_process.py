"""Python entry of the post-processing path, API-compatible with the reference's
openpose_plus/inference/post_process.py (PostProcessor :109-150, Human :39-55, BodyPart :58-79).

    pp = PostProcessor(origin_size=(H, W), feature_size=(h, w), data_format='channels_last')
    humans, heatmap_up, pafmap_up = pp(heatmap, pafmap)

`humans` are `Human` objects with normalised `BodyPart` coordinates, `heatmap_up` / `pafmap_up` are
the up-sampled maps in [H, W, C] layout, as in the reference.  The work is done by the CUDA kernels
behind the C-ABI (include/opp_b200.h); the grouping is the C++ reference's algorithm (src/paf.cpp),
which the reference's Python path delegates to an external `pafprocess` module instead.
"""
import numpy as np

from . import _capi as capi
from .engine import Engine


class BodyPart:
    """part_idx: COCO part index (0 = nose); x, y: coordinates normalised to [0, 1); score: confidence."""
    __slots__ = ('uidx', 'part_idx', 'x', 'y', 'score')

    def __init__(self, uidx, part_idx, x, y, score):
        self.uidx = uidx
        self.part_idx = part_idx
        self.x, self.y = x, y
        self.score = score

    def get_part_name(self):
        from .common import CocoPart
        return CocoPart(self.part_idx)

    def __str__(self):
        return 'BodyPart:%d-(%.2f, %.2f) score=%.2f' % (self.part_idx, self.x, self.y, self.score)

    __repr__ = __str__


class Human:
    """body_parts: {part_idx: BodyPart}"""
    __slots__ = ('body_parts', 'pairs', 'uidx_list', 'score')

    def __init__(self, pairs):
        self.pairs = pairs
        self.uidx_list = set()
        self.body_parts = {}
        self.score = 0.0

    def __str__(self):
        return ' '.join([str(x) for x in self.body_parts.values()])

    __repr__ = __str__


def humans_from_records(records, height, width):
    """human_t records (pixel coordinates of the up-sampled map) -> reference-style Human objects
    (estimate_paf, post_process.py:86-106: x / W, y / H)."""
    out = []
    for human_id, rec in enumerate(records):
        human = Human([])
        for part_idx in range(capi.N_PARTS):
            p = rec['parts'][part_idx]
            if not p['has_value']:
                continue
            human.body_parts[part_idx] = BodyPart('%d-%d' % (human_id, part_idx), part_idx, float(p['x']) / width,
                                                  float(p['y']) / height, float(p['score']))
            human.uidx_list.add('%d-%d' % (human_id, part_idx))
        if human.body_parts:
            human.score = float(rec['score'])
            out.append(human)
    return out


class PostProcessor(object):
    def __init__(self, origin_size, feature_size, data_format='channels_last', gauss_kernel_size=17, device=-1,
                 return_maps=True, maps_on_device=False, max_batch=1, variant='cpp', max_peaks_per_part=128,
                 max_cands_per_limb=1024, max_humans=128):
        """origin_size: (height, width) the maps are up-sampled to; feature_size: (height', width') of
        the feature maps; data_format: 'channels_last' ([h, w, C]) or 'channels_first' ([C, h, w]).
        variant: 'cpp' (default) = the semantics of the reference's C++ path, src/paf.cpp (the parity target);
        'python' = the semantics of the reference's own Python graph (post_process.py:13-37: CDF-derived
        kernel, zero padding -- pass gauss_kernel_size=25 for its fixed size -- and pafprocess-style grouping)."""
        if variant not in ('cpp', 'python'):
            raise ValueError("variant must be 'cpp' or 'python'")
        if data_format not in ('channels_last', 'channels_first'):
            raise ValueError('data_format must be channels_last or channels_first')
        self.data_format = data_format
        self.origin_size = tuple(origin_size)
        self.feature_size = tuple(feature_size)
        self.return_maps, self.maps_on_device = return_maps, maps_on_device
        self._engine_args = dict(gauss_kernel_size=gauss_kernel_size, max_batch=max_batch, device=device,
                                 variant=capi.VARIANT_PYTHON if variant == 'python' else capi.VARIANT_CPP)
        # starting capacities; they grow on demand (_grow) because the reference's containers are unbounded
        self._caps = dict(max_peaks_per_part=max_peaks_per_part, max_cands_per_limb=max_cands_per_limb, max_humans=max_humans)
        self.engine = self._make_engine()
        self._up = None

    def _make_engine(self):
        fs, os_ = self.feature_size, self.origin_size
        return Engine(fs[0], fs[1], os_[0], os_[1], **self._engine_args, **self._caps)

    def _grow(self, over):
        """The reference's vectors are unbounded; this build flags a frame that exceeds a fixed capacity.  A caller of
        the reference's interface must never get a truncated result: double what overflowed and run again."""
        c = self._caps
        if over & capi.FLAG_PEAK_OVERFLOW:
            c['max_peaks_per_part'] *= 2
            c['max_cands_per_limb'] *= 4
        if over & capi.FLAG_CAND_OVERFLOW:
            c['max_cands_per_limb'] *= 2
        if over & capi.FLAG_HUMAN_OVERFLOW:
            c['max_humans'] *= 2
        old = self.engine
        self.engine = self._make_engine()  # raises OppError when the device cannot hold these capacities
        old.close()

    def close(self):
        self.engine.close()

    def _maps(self, n):
        if self._up is None or self._up[0].shape[0] < n:
            import torch  # device memory plumbing only
            H, W = self.origin_size
            dev = torch.device('cuda', self.engine.device)
            self._up = (torch.empty((n, H, W, capi.N_HEAT), device=dev), torch.empty((n, H, W, capi.N_PAF), device=dev))
        return self._up[0][:n], self._up[1][:n]

    def process_batch(self, heatmaps, pafmaps):
        """[n, ...] frames -> (list of human lists, heatmap_up [n,H,W,19], pafmap_up [n,H,W,38])."""
        layout = capi.LAYOUT_HWC if self.data_format == 'channels_last' else capi.LAYOUT_CHW
        n = int(heatmaps.shape[0])
        kw = {}
        if self.return_maps:
            cu, pu = self._maps(n)
            kw = dict(conf_up=cu, paf_up=pu, up_layout=capi.LAYOUT_HWC)
        while True:
            records, counts, flags = self.engine.process(heatmaps, pafmaps, layout=layout, **kw)
            over = int(np.bitwise_or.reduce(flags)) & capi.FLAG_OVERFLOW_MASK if n else 0
            if not over:
                break
            self._grow(over)
        self.last_flags = flags
        H, W = self.origin_size
        humans = [humans_from_records(records[f, :counts[f]], H, W) for f in range(n)]
        if not self.return_maps:
            return humans, None, None
        if self.maps_on_device:
            return humans, cu, pu
        return humans, cu.cpu().numpy(), pu.cpu().numpy()

    def __call__(self, heatmap_input, pafmap_input):
        humans, hm, pm = self.process_batch(heatmap_input[None], pafmap_input[None])
        if hm is None:
            return humans[0], None, None
        return humans[0], hm[0], pm[0]

"""AnchorGenerator / ATSSAssigner / PseudoSampler / BboxOverlaps2D / DeltaXYWHBBoxCoder mirrors.

On the LD path (one square anchor of side octave_base_scale*stride per cell, `ATSSAssigner(topk=9)`,
PseudoSampler) anchors, valid flags, assignment, VLR and imitation regions are all produced by
`ldb200_ld_targets` (targets.cu) straight from the pyramid geometry; these classes validate the config
and carry its parameters (mmdet/core/anchor/anchor_generator.py:58-110, core/bbox/assigners/
atss_assigner.py:23-29, samplers/pseudo_sampler.py, iou_calculators/iou2d_calculator.py:6-40)."""
import torch

from .registry import ANCHOR_GENERATORS, BBOX_ASSIGNERS, BBOX_CODERS, BBOX_SAMPLERS, IOU_CALCULATORS


def _pair(v):
    return tuple(v) if isinstance(v, (tuple, list)) else (v, v)


@ANCHOR_GENERATORS.register_module()
class AnchorGenerator:
    def __init__(self, strides, ratios, scales=None, base_sizes=None, scale_major=True, octave_base_scale=None,
                 scales_per_octave=None, centers=None, center_offset=0.):
        if center_offset != 0:
            assert centers is None
        assert (octave_base_scale is not None and scales_per_octave is not None) ^ (scales is not None), \
            'scales and octave_base_scale with scales_per_octave cannot be set at the same time'
        self.strides = [_pair(s) for s in strides]
        self.base_sizes = [min(s) for s in self.strides] if base_sizes is None else base_sizes
        self.ratios = torch.tensor(ratios, dtype=torch.float32)
        if scales is not None:
            self.scales = torch.tensor(scales, dtype=torch.float32)
        else:
            self.scales = torch.tensor([2 ** (i / scales_per_octave) for i in range(scales_per_octave)],
                                       dtype=torch.float32) * octave_base_scale
        self.octave_base_scale, self.scales_per_octave = octave_base_scale, scales_per_octave
        self.center_offset = center_offset
        # what the sm_100a target kernel implements (LD/GFL configs): one square anchor per cell
        if not (len(ratios) == 1 and float(ratios[0]) == 1.0 and self.scales.numel() == 1 and center_offset == 0
                and all(s[0] == s[1] for s in self.strides)
                and all(b == s[0] for b, s in zip(self.base_sizes, self.strides))):
            raise NotImplementedError('ld_b200 target kernels support the GFL/LD anchor layout only '
                                      '(ratios=[1.0], one scale per octave, center_offset=0, square strides)')
        self.anchor_scale = float(self.scales[0])
        if self.anchor_scale != 8.0:
            raise NotImplementedError('ld_b200 target kernels assume octave_base_scale=8 (all LD/GFL configs)')

    @property
    def num_levels(self):
        return len(self.strides)

    @property
    def num_base_anchors(self):
        return [1 for _ in self.strides]

    def grid_anchors(self, featmap_sizes, device='cuda'):
        """anchor_generator.py:233-270; provided for API parity (the kernels never materialise anchors)."""
        out = []
        for (h, w), (s, _) in zip(featmap_sizes, self.strides):
            half = self.anchor_scale * s / 2
            xs = torch.arange(w, device=device, dtype=torch.float32) * s
            ys = torch.arange(h, device=device, dtype=torch.float32) * s
            cx, cy = xs.repeat(h), ys.repeat_interleave(w)
            out.append(torch.stack([cx - half, cy - half, cx + half, cy + half], 1))
        return out


@BBOX_ASSIGNERS.register_module()
class ATSSAssigner:
    def __init__(self, topk, iou_calculator=dict(type='BboxOverlaps2D'), ignore_iof_thr=-1):
        self.topk = topk
        self.ignore_iof_thr = ignore_iof_thr
        if ignore_iof_thr > 0:
            raise NotImplementedError('ignore_iof_thr > 0 is not used by any LD config (atss_assigner.py:105-112)')
        from .registry import build_iou_calculator
        self.iou_calculator = build_iou_calculator(iou_calculator)


@BBOX_SAMPLERS.register_module()
class PseudoSampler:
    def __init__(self, **kwargs):
        pass


@IOU_CALCULATORS.register_module()
class BboxOverlaps2D:
    pass


@BBOX_CODERS.register_module()
class DeltaXYWHBBoxCoder:
    """Constructed by AnchorHead.__init__ (anchor_head.py:77) but never called on the GFL/LD path."""

    def __init__(self, target_means=(0., 0., 0., 0.), target_stds=(1., 1., 1., 1.), clip_border=True):
        self.means, self.stds, self.clip_border = target_means, target_stds, clip_border

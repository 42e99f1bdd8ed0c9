"""Seeded synthetic inputs and weights for the LD train step.

The reference trains on COCO through mmdet's data pipeline (`mmdet/datasets/**`, out of scope);
the benchmark and the parity tests use COCO-*shaped* synthetic batches instead:
``img ~ randn(N,3,H,W)`` (already "normalised"), ``G ~ clip(Poisson(7),1,100)`` boxes per image
with continuous coordinates (no distance ties in ATSS top-k), 80 classes.

`seeded_init` fills every parameter / buffer of a detector from a generator keyed by the tensor's
*name*, so the reference model (built through the mmcv stand-in), the CPU oracle port and the
B200 build get bit-identical weights without shipping a 330 MB checkpoint.  Unlike the reference's
`init_weights` (`resnet.py:598-618` zero-inits the last BN gamma of every block, which zeroes most
backbone gradients at iteration 0) the seeded init keeps every branch alive so gradient parity is
a meaningful test.
"""
import math
import zlib

import torch


def make_img_metas(n, h, w):
    return [dict(img_shape=(h, w, 3), pad_shape=(h, w, 3), ori_shape=(h, w, 3),
                 scale_factor=1.0, flip=False) for _ in range(n)]


def make_gts(n, h, w, seed, num_classes=80, mean_boxes=7.0, max_boxes=100):
    """COCO-shaped ground truth: list of [G,4] xyxy float32 (pixels) and list of [G] int64."""
    g = torch.Generator().manual_seed(int(seed))
    boxes, labels = [], []
    lo, hi = math.log(16.0), math.log(0.6 * min(h, w))
    for _ in range(n):
        while True:
            k = int(torch.poisson(torch.tensor([mean_boxes]), generator=g).item())
            k = max(1, min(max_boxes, k))
            c = torch.rand(k, 2, generator=g) * 0.8 + 0.1
            cx, cy = c[:, 0] * w, c[:, 1] * h
            wh = torch.exp(torch.rand(k, 2, generator=g) * (hi - lo) + lo)
            x1 = (cx - wh[:, 0] / 2).clamp(0, w)
            x2 = (cx + wh[:, 0] / 2).clamp(0, w)
            y1 = (cy - wh[:, 1] / 2).clamp(0, h)
            y2 = (cy + wh[:, 1] / 2).clamp(0, h)
            b = torch.stack([x1, y1, x2, y2], 1)
            keep = ((x2 - x1) >= 4) & ((y2 - y1) >= 4)
            b = b[keep]
            lab = torch.randint(0, num_classes, (k,), generator=g)[keep]
            if b.shape[0] >= 1:
                break
        boxes.append(b.float().contiguous())
        labels.append(lab.long().contiguous())
    return boxes, labels


def make_batch(n, h, w, seed=1000, num_classes=80):
    """One synthetic batch: img [N,3,H,W] fp32 (CPU), img_metas, gt_bboxes, gt_labels."""
    g = torch.Generator().manual_seed(int(seed))
    img = torch.randn(n, 3, h, w, generator=g)
    gt_bboxes, gt_labels = make_gts(n, h, w, seed + 7919, num_classes)
    return dict(img=img, img_metas=make_img_metas(n, h, w), gt_bboxes=gt_bboxes,
                gt_labels=gt_labels)


def _gen_for(name, seed):
    return torch.Generator().manual_seed((zlib.crc32(name.encode()) ^ (seed * 2654435761)) & 0x7FFFFFFF)


@torch.no_grad()
def seeded_init(model, seed=0, prefix=''):
    """Fill all float params/buffers of `model` deterministically from their state_dict names.

    Rules (by name suffix, shared by every implementation of the detector):
      *.conv*.weight / *downsample.0.weight / gfl_* weight : N(0, sqrt(2/fan_in)) backbone+FPN,
                                                             N(0, 0.01) inside bbox_head
      bias                                                : U(-0.05, 0.05); gfl_cls.bias = -4.595
      bn*/downsample.1 weight                             : U(0.5, 1.0); last BN of a block U(0.15, 0.35)
      bn* bias / running_mean                             : U(-0.1, 0.1);  running_var U(0.8, 1.2)
      gn weight                                           : U(0.8, 1.2), gn bias U(-0.1, 0.1)
      scales.*.scale                                      : U(0.9, 1.1)
    """
    sd = model.state_dict()
    for name, t in sd.items():
        if not torch.is_floating_point(t):
            continue
        full = prefix + name
        g = _gen_for(full, seed)
        leaf = name.rsplit('.', 1)[-1]
        in_head = 'bbox_head' in full
        is_norm = any(s in name for s in ('.bn', 'bn1.', 'bn2.', 'bn3.', 'downsample.1', '.gn'))
        if name.endswith('integral.project'):
            continue
        if is_norm or name.startswith('bn') or '.gn.' in name:
            if leaf == 'weight':
                if '.gn.' in name:
                    lo, hi = 0.8, 1.2
                else:
                    # last BN of a residual block (bn3 of Bottleneck, bn2 of BasicBlock handled
                    # the same: small but non-zero so inner convs receive gradient)
                    last = name.endswith('bn3.weight')
                    lo, hi = (0.15, 0.35) if last else (0.5, 1.0)
                t.copy_(torch.rand(t.shape, generator=g) * (hi - lo) + lo)
            elif leaf == 'bias' or leaf == 'running_mean':
                t.copy_(torch.rand(t.shape, generator=g) * 0.2 - 0.1)
            elif leaf == 'running_var':
                t.copy_(torch.rand(t.shape, generator=g) * 0.4 + 0.8)
            continue
        if leaf == 'scale':
            t.copy_(torch.rand(t.shape, generator=g) * 0.2 + 0.9)
        elif leaf == 'weight' and t.dim() == 4:
            if in_head:
                std = 0.01
            else:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                # FPN convs have no ReLU after them -> gain 1; backbone convs gain sqrt(2)
                std = math.sqrt((1.0 if 'neck.' in full else 2.0) / fan_in)
            t.copy_(torch.randn(t.shape, generator=g) * std)
        elif leaf == 'bias':
            if name.endswith('gfl_cls.bias'):
                t.fill_(-4.59511985013459)
            else:
                t.copy_(torch.rand(t.shape, generator=g) * 0.1 - 0.05)
        else:
            t.copy_(torch.randn(t.shape, generator=g) * 0.02)
    return model

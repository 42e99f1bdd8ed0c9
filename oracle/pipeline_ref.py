"""CPU restatement of the reference's train pipeline (TEST INFRASTRUCTURE): Resize(keep_ratio) -> RandomFlip ->
Normalize -> Pad(size_divisor) -> DefaultFormatBundle -> collate (mmdet/datasets/pipelines/transforms.py:26-600,
formating.py) on the cv2 / numpy functions mmcv wraps (restated in oracle/mmcv_standin.py; mmcv is not vendored).
Pinned by tests/test_oracle_golden.py against the reference's own transform classes run through the stand-in
(tests/golden/p_*.pt)."""
import numpy as np
import torch

from . import mmcv_standin as mm


def synth_images(sizes, seed):
    """Seeded uint8 BGR images with smooth + noisy content, and COCO-like boxes in the original image."""
    rng = np.random.default_rng(seed)
    imgs, boxes, labels = [], [], []
    for (h, w) in sizes:
        yy, xx = np.mgrid[0:h, 0:w]
        base = np.stack([(xx * 255 / w), (yy * 255 / h), ((xx + yy) * 255 / (h + w))], -1)
        img = np.clip(base + rng.normal(0, 40, (h, w, 3)), 0, 255).astype(np.uint8)
        g = int(rng.integers(1, 6))
        x1 = rng.uniform(0, w * 0.7, g)
        y1 = rng.uniform(0, h * 0.7, g)
        bw = rng.uniform(8, w * 0.5, g)
        bh = rng.uniform(8, h * 0.5, g)
        b = np.stack([x1, y1, np.minimum(x1 + bw, w), np.minimum(y1 + bh, h)], -1).astype(np.float32)
        imgs.append(img)
        boxes.append(b)
        labels.append(rng.integers(0, 80, g))
    return imgs, boxes, labels


def train_pipeline(images, gt_bboxes, flips, img_scale=(1333, 800), mean=(123.675, 116.28, 103.53),
                   std=(58.395, 57.12, 57.375), to_rgb=True, size_divisor=32):
    """-> (batch [N,3,H,W] float32 tensor, metas, boxes) exactly as the reference pipeline + mmcv collate produce."""
    outs, metas, boxes = [], [], []
    mean, std = np.asarray(mean, dtype=np.float32), np.asarray(std, dtype=np.float32)
    for img, b, flip in zip(images, gt_bboxes, flips):
        h, w = img.shape[:2]
        r, _ = mm.imrescale(img, img_scale, return_scale=True)
        nh, nw = r.shape[:2]
        sf = np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32)
        bb = b * sf
        bb[:, 0::2] = np.clip(bb[:, 0::2], 0, nw)
        bb[:, 1::2] = np.clip(bb[:, 1::2], 0, nh)
        if flip:
            r = mm.imflip(r, 'horizontal')
            f = bb.copy()
            f[..., 0::4] = nw - bb[..., 2::4]
            f[..., 2::4] = nw - bb[..., 0::4]
            bb = f
        r = mm.imnormalize(np.ascontiguousarray(r), mean, std, to_rgb)
        r = mm.impad_to_multiple(r, size_divisor, pad_val=0)
        outs.append(torch.from_numpy(np.ascontiguousarray(r.transpose(2, 0, 1))))
        metas.append(dict(img_shape=(nh, nw, 3), pad_shape=r.shape, scale_factor=sf, flip=bool(flip)))
        boxes.append(bb)
    H, W = max(o.shape[1] for o in outs), max(o.shape[2] for o in outs)
    batch = torch.zeros(len(outs), 3, H, W)
    for i, o in enumerate(outs):       # mmcv collate: pad to the largest image of the batch (bottom / right, zeros)
        batch[i, :, :o.shape[1], :o.shape[2]] = o
    return batch, metas, boxes

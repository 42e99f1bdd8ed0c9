"""Input side of the LD train step on the GPU: the reference's train pipeline

    Resize(img_scale=(1333, 800), keep_ratio=True) -> RandomFlip(0.5) -> Normalize(mean, std, to_rgb=True) ->
    Pad(size_divisor=32) -> DefaultFormatBundle -> Collect -> collate
    (configs/ld/ld_r18_gflv1_r101_fpn_coco_1x.py:66-77; mmdet/datasets/pipelines/transforms.py:26-600)

for a whole batch: ONE kernel launch (`ldb200_preprocess_batch`) turns the decoded uint8 BGR images into the
[N, 3, H, W] fp32 network input, bit-identical to the cv2 / numpy arithmetic behind mmcv.imrescale / imflip /
imnormalize / impad; the (tiny) box arithmetic of `_resize_bboxes` / `bbox_flip` and the `img_metas` bookkeeping stay
on the host with the reference's numpy expressions.  Image decoding (LoadImageFromFile) stays outside: it is file I/O.
"""
import ctypes as C

import numpy as np
import torch

from . import lib as L


def rescale_size(old_size, scale):
    """mmcv.rescale_size (+ imrescale's rounding): old_size = (w, h), scale = (long edge, short edge) bound or a factor."""
    w, h = old_size
    if isinstance(scale, (float, int)):
        factor = float(scale)
    else:
        max_long, max_short = max(scale), min(scale)
        factor = min(max_long / max(h, w), max_short / min(h, w))
    return int(w * float(factor) + 0.5), int(h * float(factor) + 0.5)


class GpuTrainPipeline:
    """Callable batch transform with the reference transforms' constructor arguments."""

    def __init__(self, img_scale=(1333, 800), keep_ratio=True, flip_ratio=0.5, mean=(123.675, 116.28, 103.53),
                 std=(58.395, 57.12, 57.375), to_rgb=True, size_divisor=32, bbox_clip_border=True, device='cuda'):
        if not keep_ratio:
            raise NotImplementedError('keep_ratio=False (every GFL / LD config resizes with keep_ratio=True)')
        self.img_scale, self.flip_ratio = tuple(img_scale), flip_ratio
        self.mean, self.std = np.asarray(mean, dtype=np.float32), np.asarray(std, dtype=np.float32)
        self.to_rgb, self.size_divisor, self.bbox_clip_border = to_rgb, size_divisor, bbox_clip_border
        self.device = torch.device(device)

    def __call__(self, images, gt_bboxes, gt_labels, flips=None, rng=None):
        """images: list of uint8 HWC BGR numpy arrays (cv2.imread order); gt_bboxes: list of [G,4] float32 xyxy in the
        ORIGINAL image; flips: optional list of bools (else drawn with `rng` / np.random at flip_ratio).
        -> dict(img [N,3,H,W] f32 on the device, img_metas, gt_bboxes, gt_labels) ready for `model(**data)`."""
        n = len(images)
        if flips is None:
            r = rng if rng is not None else np.random
            # RandomFlip.__call__ (transforms.py:416-440): np.random.choice([direction, None], p=[ratio, 1 - ratio])
            flips = [r.choice(['horizontal', None], p=[self.flip_ratio, 1 - self.flip_ratio]) is not None for _ in range(n)]
        descs = (L.PrepImage * n)()
        metas, boxes_out, off = [], [], 0
        d = self.size_divisor
        for i, img in enumerate(images):
            assert img.dtype == np.uint8 and img.ndim == 3 and img.shape[2] == 3
            h, w = img.shape[:2]
            nw, nh = rescale_size((w, h), self.img_scale)
            # Resize._resize_img (transforms.py:203-231): per-axis factors of the rounded size
            scale_factor = np.array([nw / w, nh / h, nw / w, nh / h], dtype=np.float32)
            b = np.asarray(gt_bboxes[i], dtype=np.float32).reshape(-1, 4) * scale_factor
            if self.bbox_clip_border:
                b[:, 0::2] = np.clip(b[:, 0::2], 0, nw)
                b[:, 1::2] = np.clip(b[:, 1::2], 0, nh)
            if flips[i]:      # RandomFlip.bbox_flip (transforms.py:382-414)
                f = b.copy()
                f[..., 0::4] = nw - b[..., 2::4]
                f[..., 2::4] = nw - b[..., 0::4]
                b = f
            descs[i].src_off, descs[i].sh, descs[i].sw, descs[i].dh, descs[i].dw = off, h, w, nh, nw
            descs[i].flip = int(bool(flips[i]))
            off += (h * w * 3 + 15) // 16 * 16
            ph, pw = (nh + d - 1) // d * d, (nw + d - 1) // d * d
            metas.append(dict(ori_shape=(h, w, 3), img_shape=(nh, nw, 3), pad_shape=(ph, pw, 3), scale_factor=scale_factor,
                              flip=bool(flips[i]), flip_direction='horizontal' if flips[i] else None,
                              img_norm_cfg=dict(mean=self.mean, std=self.std, to_rgb=self.to_rgb)))
            boxes_out.append(torch.from_numpy(b))
        Hp, Wp = max(m['pad_shape'][0] for m in metas), max(m['pad_shape'][1] for m in metas)   # collate: pad to the batch max
        packed = torch.empty(off, dtype=torch.uint8).pin_memory() if self.device.type == 'cuda' else torch.empty(off, dtype=torch.uint8)
        for i, img in enumerate(images):
            o = descs[i].src_off
            packed[o:o + img.size] = torch.from_numpy(np.ascontiguousarray(img).reshape(-1))
        src = packed.to(self.device, non_blocking=True)
        dbuf = torch.frombuffer(bytearray(bytes(descs)), dtype=torch.uint8).to(self.device)
        out = torch.empty(n, 3, Hp, Wp, dtype=torch.float32, device=self.device)
        mean = (C.c_float * 3)(*self.mean.tolist())
        std = (C.c_float * 3)(*self.std.tolist())
        L.check(L.load().ldb200_preprocess_batch(L.ptr(src), L.ptr(dbuf), n, L.ptr(out), Hp, Wp, mean, std, int(self.to_rgb),
                                                 L.stream_ptr()), 'ldb200_preprocess_batch')
        return dict(img=out, img_metas=metas, gt_bboxes=[b.to(self.device) for b in boxes_out],
                    gt_labels=[torch.as_tensor(np.asarray(l)).long().to(self.device) for l in gt_labels])

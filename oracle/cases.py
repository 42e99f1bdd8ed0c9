"""Seeded parity cases shared by `oracle/make_golden.py` (which runs the UNMODIFIED reference on
them and freezes its outputs under `tests/golden/`) and by the tests (which re-create the same
inputs from the seed and compare the oracle / the CUDA path with the frozen reference outputs).
TEST INFRASTRUCTURE ONLY."""
import torch

from ld_b200 import synth

STRIDES = (8, 16, 32, 64, 128)

# (name, N, H, W, seed)
TARGET_CASES = [
    ('t_256x320_s1', 2, 256, 320, 1),
    ('t_256x320_s2', 3, 256, 320, 2),
    ('t_320x320_s3', 2, 320, 320, 3),
    ('t_800x800_s4', 2, 800, 800, 4),
    ('t_800x1344_s5', 1, 800, 1344, 5),
]

LOSS_CASES = [
    ('l_256x320_s11', 2, 256, 320, 11),
    ('l_192x192_s12', 3, 192, 192, 12),   # coarse levels get no positives -> loss_im zeroing quirk
    ('l_384x512_s13', 2, 384, 512, 13),
]

# (name, student depth, teacher depth, N, H, W, seed)
STEP_CASES = [
    ('s_r18_r18_2x128x160', 18, 18, 2, 128, 160, 21),
    ('s_r50_r101_2x256x320', 50, 101, 2, 256, 320, 22),
]
# BASELINE.json configs[0]: "ld_gflv1_r101_r50_fpn_coco_1x.py, 2 synthetic 800x800 images, CPU-only 1 iter (loss parity
# vs reference)".  Compared only in the fp32-accurate ("parity") mode of the CUDA path.
STEP_CASE_CONFIG1 = ('s_r50_r101_2x800x800', 50, 101, 2, 800, 800, 23)


def grad_projection(name, grad):
    """<grad, r_name> with a seeded standard-normal r of the gradient's shape: one number per parameter tensor that
    moves with EVERY element (a full-gradient fingerprint small enough to commit for all 175 tensors)."""
    import zlib
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7fffffff)
    r = torch.randn(grad.numel(), generator=g, dtype=torch.float64)
    return float(torch.dot(grad.detach().double().flatten().cpu(), r))


def featmap_sizes(h, w):
    """ResNet/FPN spatial arithmetic (k3 s2 p1 convs: ceil-halving) for strides 8..128."""
    def half(v):
        return (v + 1) // 2
    sizes = []
    hh, ww = half(half(half(h))), half(half(half(w)))   # stride 8
    for _ in STRIDES:
        sizes.append((hh, ww))
        hh, ww = half(hh), half(ww)
    return sizes


def loss_case_inputs(n, h, w, seed, num_classes=80, reg_ch=68, feat_ch=256):
    """Random head outputs / teacher outputs / FPN features (NCHW fp32, CPU) + gts for one case."""
    g = torch.Generator().manual_seed(seed)
    sizes = featmap_sizes(h, w)

    def rnd(c, scale, shift=0.0):
        return [torch.randn(n, c, fh, fw, generator=g) * scale + shift for fh, fw in sizes]
    d = dict(
        cls=rnd(num_classes, 1.5, -2.0), reg=rnd(reg_ch, 1.0), t_cls=rnd(num_classes, 1.5, -2.0),
        t_reg=rnd(reg_ch, 1.5), x=rnd(feat_ch, 1.0), t_x=rnd(feat_ch, 1.0))
    d['gt_bboxes'], d['gt_labels'] = synth.make_gts(n, h, w, seed + 100)
    d['img_metas'] = synth.make_img_metas(n, h, w)
    d['sizes'] = sizes
    return d


def loss_case_inputs_v2(n, h, w, seed):
    """Inputs of `LDv2Head.loss` (ld_gflv2.py:286-298) for one case: the v1 tensors with 81 class channels
    (use_sigmoid=False -> cls_out_channels = num_classes + 1), `cls` / `t_cls` reinterpreted as the student's and the
    teacher's cls_feat LOGITS, plus `score` = sigmoid(cls_feat) * quality with a random quality map in (0.2, 1): the
    reference treats joint score and cls_feat as independent inputs of the loss, so does the fixture."""
    d = loss_case_inputs(n, h, w, seed, num_classes=81)
    g = torch.Generator().manual_seed(seed + 7)
    d['cls_feat'] = d.pop('cls')
    d['score'] = [c.sigmoid() * (0.2 + 0.8 * torch.rand(c.shape[0], 1, c.shape[2], c.shape[3], generator=g))
                  for c in d['cls_feat']]
    return d


LOSS_CASES_V2 = [
    ('l2_256x320_s14', 2, 256, 320, 14),
    ('l2_192x192_s15', 3, 192, 192, 15),
]

# LDv2 full step: configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py (LDv2Head student, GFocalHead teacher, gibox, IMLoss 2.0)
# with both backbones swapped for R18.  (name, student depth, teacher depth, N, H, W, seed)
STEP_CASES_V2 = [
    ('s2_r18_r18_2x128x160', 18, 18, 2, 128, 160, 31),
]


# SURVEY.md §8d / BASELINE north_star: "student loss curve matching reference within 1e-3 over 100 synthetic iters".
# (name, student depth, teacher depth, N, H, W, seed, iters, base lr, warm-up iters): the unmodified reference model
# trained with torch.optim.SGD(momentum 0.9, wd 1e-4) under mmcv's linear warm-up formula (LrUpdaterHook,
# warmup_ratio 0.001), a fresh synthetic batch (seed 2000 + iteration) every step.
CURVE_CASE = ('c_r18_r18_2x128x160_100it', 18, 18, 2, 128, 160, 41, 100, 0.0002, 20)


# Stand-alone loss modules (registry surface): seeded inputs for QualityFocalLoss / DistributionFocalLoss /
# GIoULoss called on their own.  (name, rows, seed); the call variants are listed in LOSS_MODULE_CALLS.
LOSS_MODULE_CASES = [('m_rows300_s51', 300, 51), ('m_rows64_s52', 64, 52)]
# (tag, kwargs of forward): reduction_override / avg_factor / with or without weight
LOSS_MODULE_CALLS = [('mean', dict()), ('avg', dict(avg_factor=37.5)), ('sum', dict(reduction_override='sum')),
                     ('none', dict(reduction_override='none')), ('mean_now', dict(no_weight=True))]


def loss_module_inputs(rows, seed, num_classes=80, reg_max=16):
    g = torch.Generator().manual_seed(seed)
    d = dict(
        qfl_pred=torch.randn(rows, num_classes, generator=g) * 1.5 - 2.0,
        qfl_prob=torch.rand(rows, num_classes + 1, generator=g) * 0.9 + 0.02,       # use_sigmoid=False: 81 joint scores
        label=torch.randint(0, num_classes + 12, (rows,), generator=g).clamp(max=num_classes),   # ~13 % background
        score=torch.rand(rows, generator=g),
        weight=torch.rand(rows, generator=g) + 0.1,
        dfl_pred=torch.randn(rows, reg_max + 1, generator=g) * 2.0,
        dfl_target=torch.rand(rows, generator=g) * (reg_max - 0.1),
    )
    xy = torch.rand(rows, 2, generator=g) * 40
    wh = torch.rand(rows, 2, generator=g) * 30 + 1
    d['box_target'] = torch.cat([xy, xy + wh], 1)
    d['box_pred'] = d['box_target'] + torch.randn(rows, 4, generator=g) * 6     # some pairs end up disjoint / inverted
    return d


# Inference: `GFLHead.get_bboxes` on seeded head outputs.  (name, N, H, W, seed, rescale).  The class
# logits are as sparse as a trained detector's (about 0.5 % of the scores pass score_thr = 0.05).
INFER_CASES = [('i_256x320_s61', 2, 256, 320, 61, False), ('i_800x1344_s62', 2, 800, 1344, 62, True)]


def infer_case_inputs(n, h, w, seed, num_classes=80, reg_ch=68):
    g = torch.Generator().manual_seed(seed)
    sizes = featmap_sizes(h, w)
    cls = [torch.randn(n, num_classes, fh, fw, generator=g) * 1.2 - 6.0 for fh, fw in sizes]
    reg = [torch.randn(n, reg_ch, fh, fw, generator=g) * 1.5 for fh, fw in sizes]
    metas = synth.make_img_metas(n, h, w)
    for i, m in enumerate(metas):     # content smaller than the padded tensor, and a non-trivial rescale factor
        m['img_shape'] = (h - 8 * i - 3, w - 5 * i - 2, 3)
        m['scale_factor'] = [1.25 + 0.1 * i, 1.5, 1.25 + 0.1 * i, 1.5]
    return dict(cls=cls, reg=reg, img_metas=metas, sizes=sizes)


# Input pipeline: (name, [(h, w), ...] source sizes, seed, flips, img_scale)
PIPELINE_CASES = [('p_small_s71', [(120, 160), (97, 141), (160, 101)], 71, [True, False, True], (333, 200)),
                  ('p_coco_s72', [(427, 640), (480, 640)], 72, [False, True], (1333, 800))]

"""Tensor-level wrappers over the C-ABI (`ld_b200/lib.py`): allocate outputs with torch, pass raw
device pointers + the current CUDA stream, never synchronise."""
import ctypes as C
from dataclasses import dataclass
from typing import List, Optional

import torch

from . import lib as L

TERMS = ('loss_cls', 'loss_bbox', 'loss_dfl', 'loss_ld', 'loss_ld_vlr', 'loss_kd', 'loss_kd_neg',
         'loss_im')


def _dt(t):
    return {torch.float32: L.F32, torch.bfloat16: L.BF16}[t]


# ------------------------------------------------------------------------------------ row space
def to_rows(tensors: List[torch.Tensor]) -> torch.Tensor:
    """List of per-level NCHW tensors -> one [R, C] row-space tensor (level-major, n, y, x).
    Zero-copy per level when the tensor is channels-last in memory."""
    return torch.cat([t.permute(0, 2, 3, 1).reshape(-1, t.shape[1]) for t in tensors], 0).contiguous()


def from_rows(rows: torch.Tensor, pyr: L.Pyramid) -> List[torch.Tensor]:
    """[R, C] row-space -> list of per-level [N, C, H, W] views (channels-last memory)."""
    out, o, n = [], 0, pyr.num_imgs
    for l in range(pyr.num_levels):
        h, w = pyr.h[l], pyr.w[l]
        cnt = n * h * w
        out.append(rows[o:o + cnt].view(n, h, w, -1).permute(0, 3, 1, 2))
        o += cnt
    return out


def rows_to_levels(v: torch.Tensor, pyr: L.Pyramid) -> List[torch.Tensor]:
    """[R(,k)] row-space per-anchor array -> list of [N, A_l(,k)] (the reference's layout)."""
    out, o, n = [], 0, pyr.num_imgs
    for a in pyr.level_anchors:
        out.append(v[o:o + n * a].view(n, a, *v.shape[1:]))
        o += n * a
    return out


# ------------------------------------------------------------------------------------ targets
@dataclass
class LDTargets:
    pyr: L.Pyramid
    labels: torch.Tensor         # [R] i32
    label_weights: torch.Tensor  # [R] f32
    bbox_targets: torch.Tensor   # [R,4] f32
    vlr: torch.Tensor            # [R] f32
    im: torch.Tensor             # [R] f32
    gt_inds: torch.Tensor        # [R] i32
    sparse_rows: torch.Tensor    # [R] i32 (first counts[0] valid)
    fg_rows: torch.Tensor        # [R] i32 (first counts[1] valid)
    counts: torch.Tensor         # i32 [4 + 3*MAX_LEVELS]
    flat: Optional[torch.Tensor] = None   # the one buffer all of the above are views of (a single copy moves a target set)


def ld_targets(pyr: L.Pyramid, gt_bboxes: List[torch.Tensor], gt_labels: List[torch.Tensor],
               pad_hw, topk=9, im_mode=L.IM_FINEGRAINED, num_classes=80, device=None) -> LDTargets:
    lib = L.load()
    device = device or torch.device('cuda')
    n = pyr.num_imgs
    assert len(gt_bboxes) == len(gt_labels) == n
    offs = [0]
    for b in gt_bboxes:
        assert b.dim() == 2 and b.shape[1] == 4 and b.shape[0] >= 1, \
            'every image needs >= 1 gt box (the reference breaks on G == 0, atss_assigner.py:203-216)'
        offs.append(offs[-1] + b.shape[0])
    boxes = torch.cat([b.to(device=device, dtype=torch.float32, non_blocking=True) for b in gt_bboxes]).contiguous()
    labs = torch.cat([b.to(device=device, dtype=torch.int32, non_blocking=True) for b in gt_labels]).contiguous()
    host = torch.tensor(offs + [int(v) for hw in pad_hw for v in hw[:2]], dtype=torch.int32)
    if device.type == 'cuda':
        host = host.pin_memory()
    meta = host.to(device, non_blocking=True)
    offsets, pad = meta[:n + 1], meta[n + 1:]
    R = pyr.rows
    # one allocation: [bbox_targets 4R | labels R | label_weights R | vlr R | im R | gt_inds R | sparse R | fg R | counts]
    nc = 4 + 3 * L.MAX_LEVELS
    flat = torch.empty(11 * R + nc, dtype=torch.int32, device=device)
    seg = lambda i, k=1: flat[(3 + i) * R:(3 + i + k) * R]   # noqa: E731  (after the 4R bbox_targets block)
    f = lambda v: v.view(torch.float32)                      # noqa: E731
    t = LDTargets(pyr, seg(1), f(seg(2)), f(flat[:4 * R]).view(R, 4), f(seg(3)), f(seg(4)), seg(5), seg(6), seg(7),
                  flat[11 * R:11 * R + nc], flat)
    ws = torch.empty(lib.ldb200_targets_workspace_bytes(C.byref(pyr), offs[-1]), dtype=torch.uint8,
                     device=device)
    L.check(lib.ldb200_ld_targets(C.byref(pyr), L.ptr(boxes), L.ptr(labs), L.ptr(offsets), offs[-1],
                                  L.ptr(pad), topk, im_mode, num_classes, L.ptr(t.labels),
                                  L.ptr(t.label_weights), L.ptr(t.bbox_targets), L.ptr(t.vlr), L.ptr(t.im),
                                  L.ptr(t.gt_inds), L.ptr(t.sparse_rows), L.ptr(t.fg_rows),
                                  L.ptr(t.counts), L.ptr(ws), L.stream_ptr()), 'ldb200_ld_targets')
    return t


# ------------------------------------------------------------------------------------ loss
def make_loss_cfg(num_classes=80, reg_max=16, feat_channels=256, qfl_weight=1.0, qfl_beta=2.0,
                  giou_weight=2.0, giou_eps=1e-6, dfl_weight=0.25, ld_weight=0.25, ld_T=10.0,
                  ld_vlr_weight=0.25, ld_vlr_T=10.0, kd_weight=10.0, kd_T=2.0, im_weight=2.0) -> L.LossCfg:
    return L.LossCfg(num_classes, reg_max, feat_channels, qfl_weight, qfl_beta, giou_weight, giou_eps,
                     dfl_weight, ld_weight, ld_T, ld_vlr_weight, ld_vlr_T, kd_weight, kd_T, im_weight)


def ld_loss_prepass(cfg: L.LossCfg, cls, reg, tg: LDTargets, want_score=False):
    """-> (wsum [1] f32 device, score [R] f32 or None)."""
    lib = L.load()
    wsum = torch.empty(1, dtype=torch.float32, device=cls.device)
    score = torch.empty(tg.pyr.rows, dtype=torch.float32, device=cls.device) if want_score else None
    L.check(lib.ldb200_ld_loss_prepass(C.byref(tg.pyr), C.byref(cfg), L.ptr(cls), L.ptr(reg),
                                       L.ptr(tg.labels), L.ptr(tg.bbox_targets), L.ptr(tg.sparse_rows),
                                       L.ptr(tg.counts), L.ptr(score), L.ptr(wsum), L.stream_ptr()),
            'ldb200_ld_loss_prepass')
    return wsum, score


def loss_normalisers(tg: LDTargets, wsum: torch.Tensor, group=None):
    """[num_total_samples, avg_factor] on the device (ld_head.py:338-341, 362-363); one 8-byte
    all-reduce when distributed, no host sync."""
    norm = torch.stack([tg.counts[2].float(), wsum[0] + 1e-6])
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(norm, group=group)
        norm = norm / dist.get_world_size(group)
    norm[0].clamp_(min=1.0)
    return norm


def ld_loss_fused(cfg: L.LossCfg, cls, reg, t_cls, t_reg, x, t_x, tg: LDTargets, norm,
                  term_grad: Optional[torch.Tensor] = None, grad_dtype=torch.float32, max_fg=None,
                  d_cls_ld=0, d_reg_ld=0):
    """One launch: losses [9, MAX_LEVELS] (+ d_cls [R,C], d_reg [R,4*bins], d_im [max_fg,F] when
    `term_grad` ([8, MAX_LEVELS] host floats or True for all-ones) is given)."""
    lib = L.load()
    dev = cls.device
    R = tg.pyr.rows
    assert cls.shape == (R, cfg.num_classes) and reg.shape == (R, 4 * (cfg.reg_max + 1))
    assert x.dtype == t_x.dtype and x.shape == t_x.shape == (R, cfg.feat_channels)
    losses = torch.empty(L.NUM_TERMS, L.MAX_LEVELS, dtype=torch.float32, device=dev)
    d_cls = d_reg = d_im = None
    tgp = None
    if term_grad is not None:
        if term_grad is True:
            term_grad = torch.ones(L.NUM_TERMS - 1, L.MAX_LEVELS)
        tgh = term_grad.detach().to('cpu', torch.float32).contiguous()
        assert tgh.shape == (L.NUM_TERMS - 1, L.MAX_LEVELS)
        tgp = C.c_void_p(tgh.data_ptr())
        max_fg = R if max_fg is None else max_fg
        d_cls = torch.empty(R, d_cls_ld or cfg.num_classes, dtype=grad_dtype, device=dev)
        d_reg = torch.empty(R, d_reg_ld or 4 * (cfg.reg_max + 1), dtype=grad_dtype, device=dev)
        d_im = torch.empty(max_fg, cfg.feat_channels, dtype=grad_dtype, device=dev)
    L.check(lib.ldb200_ld_loss_fused(C.byref(tg.pyr), C.byref(cfg), L.ptr(cls), L.ptr(reg), L.ptr(t_cls),
                                     L.ptr(t_reg), L.ptr(x), L.ptr(t_x), _dt(x.dtype), L.ptr(tg.labels),
                                     L.ptr(tg.label_weights), L.ptr(tg.bbox_targets), L.ptr(tg.vlr), None,
                                     L.ptr(tg.sparse_rows), L.ptr(tg.fg_rows), L.ptr(tg.counts),
                                     L.ptr(norm), tgp, L.ptr(losses), L.ptr(d_cls), L.ptr(d_reg),
                                     L.ptr(d_im), int(d_cls_ld), int(d_reg_ld), _dt(grad_dtype), L.stream_ptr()),
            'ldb200_ld_loss_fused')
    return losses, d_cls, d_reg, d_im


def gi_region(cfg: L.LossCfg, cls, reg, t_cls, t_reg, tg: LDTargets, iou_thr=0.3, topk=10):
    """`LDHead.get_gi_region` for all levels (ldb200_gi_region): -> (gi_rows i32 [L*topk], counts i32) to be used
    as (fg_rows, counts) of the fused loss when imitation_method='gibox'."""
    lib = L.load()
    dev = cls.device
    nl = tg.pyr.num_levels
    ws = torch.empty(int(lib.ldb200_gi_region_workspace_bytes(C.byref(tg.pyr), int(topk))), dtype=torch.uint8, device=dev)
    gi_rows = torch.empty(nl * topk, dtype=torch.int32, device=dev)
    counts = torch.empty_like(tg.counts)
    for t in (cls, reg, t_cls, t_reg):
        assert t.dtype == torch.float32 and t.is_contiguous()
    L.check(lib.ldb200_gi_region(C.byref(tg.pyr), C.byref(cfg), L.ptr(cls), L.ptr(reg), L.ptr(t_cls), L.ptr(t_reg),
                                 float(iou_thr), int(topk), L.ptr(tg.counts), L.ptr(gi_rows), L.ptr(counts), L.ptr(ws),
                                 L.stream_ptr()), 'ldb200_gi_region')
    return gi_rows, counts


def dgqp_forward(reg, cls_feat, W1, b1, W2, b2, reg_topk=4, add_mean=True):
    """GFLv2 quality branch + joint score on row-space head outputs (ldb200_dgqp_forward, gfocal_head.py:198-215).
    reg [R, 4*bins] f32, cls_feat [R, C] f32; W1 [hid, 4*(topk+add_mean)], b1 [hid], W2 [hid], b2 [1] (fp32, device).
    -> (quality [R], score [R, C])."""
    lib = L.load()
    R, C_ = cls_feat.shape
    bins = reg.shape[1] // 4
    for t in (reg, cls_feat, W1, b1, W2, b2):
        assert t.dtype == torch.float32 and t.is_contiguous()
    hid = W1.shape[0]
    assert W1.shape[1] == 4 * (reg_topk + int(add_mean)) and W2.numel() == hid and b1.numel() == hid and b2.numel() == 1
    q = torch.empty(R, dtype=torch.float32, device=reg.device)
    score = torch.empty(R, C_, dtype=torch.float32, device=reg.device)
    L.check(lib.ldb200_dgqp_forward(L.ptr(reg), 0, L.ptr(cls_feat), 0, R, C_, bins, int(reg_topk), int(add_mean), hid,
                                    L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), L.ptr(q), L.ptr(score), L.stream_ptr()),
            'ldb200_dgqp_forward')
    return q, score


def dgqp_backward(reg, cls_feat, W1, b1, W2, b2, quality, d_score, d_reg, d_kd=None, reg_topk=4, add_mean=True):
    """Backward of `dgqp_forward`: -> (d_cls_feat, dW1, db1, dW2, db2); the predictor's gradient w.r.t. the box
    distribution logits is ADDED into `d_reg` in place.  d_score / d_kd / d_reg share one dtype (f32 or bf16)."""
    lib = L.load()
    R, C_ = cls_feat.shape
    bins = reg.shape[1] // 4
    hid = W1.shape[0]
    assert d_score.shape[0] == R and d_score.stride(1) == 1 and d_reg.stride(1) == 1 and d_reg.dtype == d_score.dtype
    assert d_kd is None or (d_kd.dtype == d_score.dtype and d_kd.stride() == d_score.stride())
    d_cls_feat = torch.empty_like(d_score)
    dW1, db1, dW2, db2 = (torch.zeros_like(W1), torch.zeros_like(b1), torch.zeros_like(W2), torch.zeros_like(b2))
    L.check(lib.ldb200_dgqp_backward(L.ptr(reg), 0, L.ptr(cls_feat), 0, R, C_, bins, int(reg_topk), int(add_mean), hid,
                                     L.ptr(W1), L.ptr(b1), L.ptr(W2), L.ptr(b2), L.ptr(quality), L.ptr(d_score),
                                     L.ptr(d_kd), int(d_score.stride(0)), L.ptr(d_cls_feat), L.ptr(d_reg),
                                     int(d_reg.stride(0)), _dt(d_score.dtype), L.ptr(dW1), L.ptr(db1), L.ptr(dW2),
                                     L.ptr(db2), L.stream_ptr()), 'ldb200_dgqp_backward')
    return d_cls_feat, dW1, db1, dW2, db2


# ---- LDv2Head (ld_gflv2.py) -------------------------------------------------------------
def ld_loss_prepass_v2(cfg: L.LossCfg, score, reg, tg: LDTargets, num_fg, want_score=False):
    """`ld_loss_prepass` on joint scores (weight_targets = max_c score, ld_gflv2.py:200)."""
    lib = L.load()
    wsum = torch.empty(1, dtype=torch.float32, device=score.device)
    iou = torch.empty(tg.pyr.rows, dtype=torch.float32, device=score.device) if want_score else None
    L.check(lib.ldb200_ld_loss_prepass_v2(C.byref(tg.pyr), C.byref(cfg), L.ptr(score), L.ptr(reg), L.ptr(tg.labels),
                                          L.ptr(tg.bbox_targets), L.ptr(tg.sparse_rows), L.ptr(tg.counts), int(num_fg),
                                          L.ptr(iou), L.ptr(wsum), L.stream_ptr()), 'ldb200_ld_loss_prepass_v2')
    return wsum, iou


def ld_loss_fused_v2(cfg: L.LossCfg, score, cls_feat, reg, t_cls_feat, t_reg, x, t_x, tg: LDTargets, norm, num_fg,
                     term_grad=None, grad_dtype=torch.float32, max_fg=None, d_cls_ld=0, d_reg_ld=0):
    """`LDv2Head.loss_single` x L in one launch: -> (losses, d_score, d_kd, d_reg, d_im); d_score is the gradient
    w.r.t. the joint score, d_kd the KD term's gradient w.r.t. cls_feat (both feed `dgqp_backward`)."""
    lib = L.load()
    dev = score.device
    R = tg.pyr.rows
    assert score.shape == cls_feat.shape == t_cls_feat.shape == (R, cfg.num_classes)
    assert reg.shape == t_reg.shape == (R, 4 * (cfg.reg_max + 1))
    assert x.dtype == t_x.dtype and x.shape == t_x.shape == (R, cfg.feat_channels)
    losses = torch.empty(L.NUM_TERMS, L.MAX_LEVELS, dtype=torch.float32, device=dev)
    d_score = d_kd = d_reg = d_im = None
    tgp = None
    if term_grad is not None:
        if term_grad is True:
            term_grad = torch.ones(L.NUM_TERMS - 1, L.MAX_LEVELS)
        tgh = term_grad.detach().to('cpu', torch.float32).contiguous()
        tgp = C.c_void_p(tgh.data_ptr())
        max_fg = R if max_fg is None else max_fg
        ld = d_cls_ld or (cfg.num_classes + 3) // 4 * 4
        d_cls_ld = ld
        d_score = torch.zeros(R, ld, dtype=grad_dtype, device=dev)    # padding columns stay zero
        d_kd = torch.empty(R, ld, dtype=grad_dtype, device=dev)
        d_reg = torch.empty(R, d_reg_ld or 4 * (cfg.reg_max + 1), dtype=grad_dtype, device=dev)
        d_im = torch.empty(max_fg, cfg.feat_channels, dtype=grad_dtype, device=dev)
    L.check(lib.ldb200_ld_loss_fused_v2(C.byref(tg.pyr), C.byref(cfg), L.ptr(score), L.ptr(cls_feat), L.ptr(reg),
                                        L.ptr(t_cls_feat), L.ptr(t_reg), L.ptr(x), L.ptr(t_x), _dt(x.dtype),
                                        L.ptr(tg.labels), L.ptr(tg.label_weights), L.ptr(tg.bbox_targets), L.ptr(tg.vlr),
                                        L.ptr(tg.sparse_rows), L.ptr(tg.fg_rows), L.ptr(tg.counts), L.ptr(norm), tgp,
                                        int(num_fg), L.ptr(losses), L.ptr(d_score), L.ptr(d_kd), L.ptr(d_reg), L.ptr(d_im),
                                        int(d_cls_ld), int(d_reg_ld), _dt(grad_dtype), L.stream_ptr()),
            'ldb200_ld_loss_fused_v2')
    return losses, d_score, d_kd, d_reg, d_im


def gi_region_v2(cfg: L.LossCfg, score, reg, t_cls_feat, t_reg, tg: LDTargets, iou_thr=0.3, topk=10):
    """`LDv2Head.get_gi_region`: GI score |teacher cls_feat - student joint score|, no sigmoid (ldb200_gi_region_v2)."""
    lib = L.load()
    dev = score.device
    ws = torch.empty(int(lib.ldb200_gi_region_workspace_bytes(C.byref(tg.pyr), int(topk))), dtype=torch.uint8, device=dev)
    gi_rows = torch.empty(tg.pyr.num_levels * topk, dtype=torch.int32, device=dev)
    counts = torch.empty_like(tg.counts)
    L.check(lib.ldb200_gi_region_v2(C.byref(tg.pyr), C.byref(cfg), L.ptr(score), L.ptr(reg), L.ptr(t_cls_feat),
                                    L.ptr(t_reg), float(iou_thr), int(topk), L.ptr(tg.counts), L.ptr(gi_rows),
                                    L.ptr(counts), L.ptr(ws), L.stream_ptr()), 'ldb200_gi_region_v2')
    return gi_rows, counts


def gfl_get_bboxes(pyr: L.Pyramid, cls_rows, reg_rows, img_hw, scale_factors=None, nms_pre=1000, score_thr=0.05,
                   iou_thr=0.6, max_per_img=100, use_sigmoid=True):
    """ldb200_gfl_get_bboxes: -> (dets [N, max_per_img, 5] f32, labels [N, max_per_img] i32, counts [N] i32)."""
    lib = L.load()
    dev = cls_rows.device
    n = pyr.num_imgs
    C_, bins = cls_rows.shape[1], reg_rows.shape[1] // 4
    assert cls_rows.dtype == reg_rows.dtype == torch.float32 and cls_rows.is_contiguous() and reg_rows.is_contiguous()
    hw = torch.tensor([[int(s[0]), int(s[1])] for s in img_hw], dtype=torch.int32).to(dev)
    sf = None if scale_factors is None else torch.tensor([list(map(float, f)) for f in scale_factors],
                                                         dtype=torch.float32).to(dev)
    ws = torch.empty(int(lib.ldb200_get_bboxes_workspace_bytes(C.byref(pyr), C_, int(nms_pre))), dtype=torch.uint8, device=dev)
    dets = torch.empty(n, max_per_img, 5, dtype=torch.float32, device=dev)
    labels = torch.empty(n, max_per_img, dtype=torch.int32, device=dev)
    counts = torch.empty(n, dtype=torch.int32, device=dev)
    L.check(lib.ldb200_gfl_get_bboxes(C.byref(pyr), L.ptr(cls_rows), L.ptr(reg_rows), C_, bins, int(use_sigmoid), L.ptr(hw), L.ptr(sf),
                                      int(nms_pre), float(score_thr), float(iou_thr), int(max_per_img), L.ptr(dets),
                                      L.ptr(labels), L.ptr(counts), L.ptr(ws), L.stream_ptr()), 'ldb200_gfl_get_bboxes')
    return dets, labels, counts


def scatter_add_rows(dst, src, rows, n_rows_dev, max_rows=None):
    lib = L.load()
    max_rows = src.shape[0] if max_rows is None else max_rows
    L.check(lib.ldb200_scatter_add_rows(L.ptr(dst), L.ptr(src), L.ptr(rows), L.ptr(n_rows_dev),
                                        int(max_rows), int(dst.shape[-1]), _dt(dst.dtype),
                                        L.stream_ptr()), 'ldb200_scatter_add_rows')


def kl_dense_fwd_bwd(student, teacher, weight, T=10.0, loss_weight=0.25, avg_factor=4.0, grad_out=1.0,
                     loss=None, d_student=None):
    """Dense LD-KL forward+backward (BASELINE configs[4]); accumulates into `loss` [1]."""
    lib = L.load()
    rows, bins = student.shape
    if loss is None:
        loss = torch.zeros(1, dtype=torch.float32, device=student.device)
    if d_student is None:
        d_student = torch.empty_like(student)
    L.check(lib.ldb200_kl_dense_fwd_bwd(L.ptr(student), L.ptr(teacher), L.ptr(weight), rows, bins,
                                        float(T), float(loss_weight), float(avg_factor), float(grad_out),
                                        L.ptr(loss), L.ptr(d_student), L.stream_ptr()),
            'ldb200_kl_dense_fwd_bwd')
    return loss, d_student


# ------------------------------------------------------------------------------------ convolution
def pack_conv_weight(w: torch.Tensor) -> torch.Tensor:
    """[Cout, Cin, kh, kw] (PyTorch OIHW) -> [Cout, kh*kw*Cin] bf16, tap-major then channel
    (the K-major operand layout of ldb200_conv2d_tc)."""
    co, ci, kh, kw = w.shape
    return w.permute(0, 2, 3, 1).reshape(co, kh * kw * ci).to(torch.bfloat16).contiguous()


def pack_conv_weight_dgrad(w: torch.Tensor) -> torch.Tensor:
    """Weights of the data-gradient conv: roles of Cin/Cout swapped, taps kept in forward order
    (the tap table of `conv2d_tc` is given the mirrored offsets)."""
    return pack_conv_weight(w.transpose(0, 1))


PARITY_PLANES = 3     # bf16 planes per fp32 operand in the fp32-accurate mode (3: ~2^-23, 2: ~2^-16 per product)


def split_bf16(x, planes=None, ld_out=None):
    """fp32 [..., C] (last dim contiguous, uniform row stride) -> [planes, rows, ld_out] bf16 with x ~= sum of planes
    (ldb200_split_bf16): the operand format of the split-bf16 ("parity") convolutions."""
    planes = planes or PARITY_PLANES
    assert x.dtype == torch.float32 and x.stride(-1) == 1
    c = x.shape[-1]
    ld_in = x.stride(-2) if x.dim() > 1 else c
    rows = x.numel() // c
    # rows must be uniformly strided (a channel slice of a contiguous [..., ld] tensor)
    chk = ld_in
    for i in range(x.dim() - 2, -1, -1):
        assert x.shape[i] == 1 or x.stride(i) == chk, 'split_bf16: rows must be uniformly strided'
        chk *= x.shape[i]
    ld_out = ld_out or (c + 7) // 8 * 8
    out = torch.empty(planes, rows, ld_out, dtype=torch.bfloat16, device=x.device)
    L.check(L.load().ldb200_split_bf16(C.c_void_p(x.data_ptr()), rows, c, ld_in, planes, L.ptr(out), ld_out,
                                       L.stream_ptr()), 'ldb200_split_bf16')
    return out


def split_weight(w2d, planes=None):
    """[Cout, K] fp32 -> [Cout, planes, K8] bf16 (K8 = K rounded up to 8): the weight operand of a split-bf16 conv."""
    planes = planes or PARITY_PLANES
    co, k = w2d.shape
    sp = split_bf16(w2d.contiguous().float(), planes)        # [planes, Cout, K8]
    return sp.permute(1, 0, 2).contiguous()


# fp32-accurate mode: leading plane product and corrections in two accumulation chains (LDB200_PARITY_ONE_PASS=1: all six
# plane products in one chain, for A/B measurements of the accumulation error)
import os as _os
PARITY_TWO_PASS = _os.environ.get('LDB200_PARITY_ONE_PASS') is None


# Tensor cores accumulate with truncation (round toward zero): a chain of k MMA steps into one TMEM accumulator loses
# about k/2 ulp of the running sum, systematically.  The fp32-accurate mode therefore keeps every chain of the LEADING
# plane product short -- one filter tap and at most PARITY_CHAIN_CH input channels (256 -> 16 MMA steps) -- and adds the
# partial results in the epilogue with round-to-nearest fp32 adds (`pre_src`); the correction products are 2^-8 of the
# leading one, their (long) chain contributes 2^-8 of the error.  LDB200_PARITY_CHAIN_CH=0: one chain for the leading term.
PARITY_CHAIN_CH = int(_os.environ.get('LDB200_PARITY_CHAIN_CH', '256'))


def _chain_descs(d1):
    """Sub-descriptors of a leading-product pass: one launch per (tap, <= PARITY_CHAIN_CH input channels)."""
    if PARITY_CHAIN_CH <= 0:
        return [d1]
    out = []
    cin = d1.Cin
    for t in range(d1.ntaps):
        for c0 in range(0, cin, PARITY_CHAIN_CH):
            d = L.ConvDesc.from_buffer_copy(d1)
            d.ntaps = 1
            d.tap_dx[0], d.tap_dy[0], d.tap_wk[0] = d1.tap_dx[t], d1.tap_dy[t], d1.tap_wk[t] + c0
            d.in_ = d1.in_ + 2 * c0
            d.Cin = min(PARITY_CHAIN_CH, cin - c0)
            out.append(d)
    return out


def _conv_desc(d, x, w_packed, cout, kh, kw, stride, pad, out, out_dtype, scale, bias, add, mask, relu, accumulate,
               flat, taps, out_hw, out_map, xsplit=None):
    assert x.dim() == 4
    keep = None
    if x.dtype == torch.float32:
        # fp32-accurate mode: split the activation into bf16 planes, weights arrive as [Cout, planes, K8]
        assert w_packed.dim() == 3, 'fp32 activations need split weights (ops.split_weight)'
        planes = w_packed.shape[1]
        n_, h_, w__, c_ = x.shape
        keep = xsplit if xsplit is not None else split_bf16(x, planes)
        x = keep[0].view(n_, h_, w__, -1)[..., :c_]
        d.split, d.w_plane = planes, w_packed.shape[2]
        out_dtype = torch.float32
        w_packed = w_packed.view(cout, -1)
    assert x.dtype == torch.bfloat16
    n, h, w_, cin = x.shape
    in_ld = x.stride(2)
    assert x.stride(3) == 1 and x.stride(1) == w_ * in_ld and x.stride(0) == h * w_ * in_ld
    if out_hw is None:
        ho = (h + 2 * pad - kh) // stride + 1
        wo = (w_ + 2 * pad - kw) // stride + 1
    else:
        ho, wo = out_hw
    d.in_ = x.data_ptr()
    d.N, d.H, d.W, d.Cin, d.in_ld = n, h, w_, cin, in_ld
    d.weight = w_packed.data_ptr()
    d.Cout, d.w_k = cout, w_packed.shape[1]
    if taps is None:
        taps = [(s - pad, r - pad, (r * kw + s) * cin) for r in range(kh) for s in range(kw)]
    d.ntaps = len(taps)
    for i, (dx, dy, wk) in enumerate(taps):
        d.tap_dx[i], d.tap_dy[i], d.tap_wk[i] = dx, dy, wk
    d.sx = d.sy = stride
    d.HO, d.WO = ho, wo
    if flat is None:
        flat = (kh == 1 and kw == 1 and stride == 1 and pad == 0 and taps == [(0, 0, 0)] and out_map is None
                and (add is None or tuple(add.shape[1:3]) == (ho, wo)))
    d.flat = int(flat)
    if out is None:
        oh, ow = (ho, wo) if out_map is None else out_map[0]
        alloc = torch.zeros if out_map is not None else torch.empty
        out = alloc(n, oh, ow, cout, dtype=out_dtype, device=x.device)
    assert out.stride(3) == 1
    d.out = out.data_ptr()
    d.OH, d.OW = out.shape[1], out.shape[2]
    if out_map is not None:
        (_, _), (d.oys, d.oyo, d.oxs, d.oxo) = out_map
    d.out_ld = out.stride(2)
    d.out_f32 = int(out.dtype == torch.float32)
    for name, t in (('scale', scale), ('bias', bias)):
        if t is not None:
            assert t.dtype == torch.float32 and t.numel() == cout and t.is_contiguous()
            setattr(d, name, t.data_ptr())
    if add is not None:
        assert add.dtype in (torch.bfloat16, torch.float32) and add.stride(3) == 1
        d.add_src, d.add_H, d.add_W, d.add_ld = add.data_ptr(), add.shape[1], add.shape[2], add.stride(2)
        d.add_f32 = int(add.dtype == torch.float32)
    if mask is not None:
        assert mask.dtype in (torch.bfloat16, torch.float32) and mask.shape[:3] == out.shape[:3]
        d.mask_src, d.mask_ld = mask.data_ptr(), mask.stride(2)
        d.mask_f32 = int(mask.dtype == torch.float32)
    d.relu, d.accumulate = int(relu), int(accumulate)
    return out, keep    # the split planes must stay referenced until the launch is enqueued (stream-ordered allocator)


def conv2d_tc(x, w_packed, cout, kh=1, kw=1, stride=1, pad=0, out=None, out_dtype=torch.bfloat16,
              scale=None, bias=None, add=None, mask=None, relu=False, accumulate=False, flat=None,
              taps=None, out_hw=None, out_map=None):
    """NHWC bf16 convolution on tcgen05 (ldb200_conv2d_tc).  x: [N,H,W,Cin] bf16 contiguous (or a
    channel slice with stride(2) = ld).  Returns [N,HO,WO,Cout] NHWC `out_dtype`."""
    if x.dtype == torch.float32 and PARITY_TWO_PASS:
        # pass 1: leading plane product alone (plain fp32 store into a scratch tensor with the output's geometry);
        # pass 2: the correction products + the result of pass 1 (pre_src), then the whole epilogue
        d1, d2 = L.ConvDesc(), L.ConvDesc()
        tmp_like = None
        if out is not None:
            tmp_like = torch.empty(out.shape, dtype=torch.float32, device=x.device)
        t1, keep = _conv_desc(d1, x, w_packed, cout, kh, kw, stride, pad, tmp_like, torch.float32, None, None, None, None,
                              False, False, flat, taps, out_hw, out_map)
        d1.split_pass = 1
        for i, dd in enumerate(_chain_descs(d1)):     # short chains, partial sums added in place (same thread reads
            if i:                                     # and writes an element, so pre_src == out is safe)
                dd.pre_src = t1.data_ptr()
            L.check(L.load().ldb200_conv2d_tc(C.byref(dd), L.stream_ptr()), 'ldb200_conv2d_tc')
        out, _ = _conv_desc(d2, x, w_packed, cout, kh, kw, stride, pad, out, out_dtype, scale, bias, add, mask, relu,
                            accumulate, flat, taps, out_hw, out_map, xsplit=keep)
        assert out.stride(2) == t1.stride(2), 'two-pass split conv: scratch and output rows must have the same stride'
        d2.split_pass, d2.pre_src = 2, t1.data_ptr()
        L.check(L.load().ldb200_conv2d_tc(C.byref(d2), L.stream_ptr()), 'ldb200_conv2d_tc')
        del keep, t1
        return out
    d = L.ConvDesc()
    out, keep = _conv_desc(d, x, w_packed, cout, kh, kw, stride, pad, out, out_dtype, scale, bias, add, mask, relu,
                           accumulate, flat, taps, out_hw, out_map)
    L.check(L.load().ldb200_conv2d_tc(C.byref(d), L.stream_ptr()), 'ldb200_conv2d_tc')
    del keep
    return out


def conv2d_tc_multi(xs, w_packed, cout, kh=1, kw=1, stride=1, pad=0, outs=None, out_dtype=torch.bfloat16,
                    scale=None, bias=None, masks=None, relu=False, taps=None, gn_stats=None):
    """The same convolution on several inputs (the FPN levels of a level-shared head conv) in ONE launch.
    gn_stats: optional list of zero-filled [N, cout/8, 2] f32 tensors the epilogue accumulates the GroupNorm
    statistics of each output into."""
    n = len(xs)
    pre, xsplits = None, [None] * n
    if xs[0].dtype == torch.float32 and PARITY_TWO_PASS:
        d1 = (L.ConvDesc * n)()
        pre, xsplits = [], []
        for i, x in enumerate(xs):
            tmp = None if outs is None else torch.empty(outs[i].shape, dtype=torch.float32, device=x.device)
            t1, k = _conv_desc(d1[i], x, w_packed, cout, kh, kw, stride, pad, tmp, torch.float32, None, None, None, None,
                               False, False, False, taps, None if taps is None else tuple(x.shape[1:3]), None)
            d1[i].split_pass = 1
            pre.append(t1)
            xsplits.append(k)
        chains = [_chain_descs(d1[i]) for i in range(n)]
        for j in range(len(chains[0])):
            dj = (L.ConvDesc * n)(*[chains[i][j] for i in range(n)])
            if j:
                for i in range(n):
                    dj[i].pre_src = pre[i].data_ptr()
            L.check(L.load().ldb200_conv2d_tc_multi(dj, n, L.stream_ptr()), 'ldb200_conv2d_tc_multi')
    descs = (L.ConvDesc * n)()
    res, keep = [], []
    for i, x in enumerate(xs):
        o, k = _conv_desc(descs[i], x, w_packed, cout, kh, kw, stride, pad, None if outs is None else outs[i],
                          out_dtype, scale, bias, None, None if masks is None else masks[i], relu, False, False,
                          taps, None if taps is None else tuple(x.shape[1:3]), None, xsplit=xsplits[i])
        if pre is not None:
            assert o.stride(2) == pre[i].stride(2)
            descs[i].split_pass, descs[i].pre_src = 2, pre[i].data_ptr()
        if gn_stats is not None:
            assert gn_stats[i].dtype == torch.float32 and gn_stats[i].numel() == x.shape[0] * (cout // 8) * 2
            descs[i].gn_stats = gn_stats[i].data_ptr()
        res.append(o)
        keep.append(k)
    L.check(L.load().ldb200_conv2d_tc_multi(descs, n, L.stream_ptr()), 'ldb200_conv2d_tc_multi')
    del keep, pre, xsplits
    return res



# ------------------------------------------------------------------------------------ glue kernels
def stem_im2col(img, dtype=torch.bfloat16):
    """img [N,3,H,W] f32 NCHW -> ([N*HO*WO, 192] bf16 (or f32: parity mode), HO, WO)."""
    n, c, h, w = img.shape
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    out = torch.empty(n * ho * wo, 192, dtype=dtype, device=img.device)
    L.check(L.load().ldb200_stem_im2col(L.ptr(img), L.ptr(out), n, h, w, _dt(dtype), L.stream_ptr()), 'ldb200_stem_im2col')
    return out, ho, wo


def stem_tc(img, weights, affines):
    """Fused conv1 7x7/2 + folded BN + ReLU for 1 or 2 networks sharing `img` (ldb200_stem_tc).
    img [N,3,H,W] f32; weights [64*nets, 192] bf16; affines = [(scale, shift), ...] per network -> list of
    [N,HO,WO,64] bf16 NHWC tensors."""
    n, c, h, w = img.shape
    nets = len(affines)
    assert c == 3 and img.dtype == torch.float32 and img.is_contiguous() and nets in (1, 2)
    assert weights.shape == (64 * nets, 192) and weights.dtype == torch.bfloat16 and weights.is_contiguous()
    ho, wo = (h - 1) // 2 + 1, (w - 1) // 2 + 1
    outs = [torch.empty(n, ho, wo, 64, dtype=torch.bfloat16, device=img.device) for _ in range(nets)]
    sc = [a[0].detach().float().contiguous() for a in affines]
    sh = [a[1].detach().float().contiguous() for a in affines]
    L.check(L.load().ldb200_stem_tc(L.ptr(img), n, h, w, L.ptr(weights), nets, L.ptr(sc[0]), L.ptr(sh[0]), L.ptr(outs[0]),
                                    L.ptr(sc[1]) if nets == 2 else None, L.ptr(sh[1]) if nets == 2 else None,
                                    L.ptr(outs[1]) if nets == 2 else None, L.stream_ptr()), 'ldb200_stem_tc')
    return outs


def deform_im2col(x, offset, kh=3, kw=3, stride=1, pad=1, dilation=1, deform_groups=1):
    """Bilinear-sampled im2col of DCNv1 (ldb200_deform_im2col): x [N,H,W,C], offset [N,HO,WO,2*dg*kh*kw] f32 ->
    [N*HO*WO, kh*kw*C] in x's dtype (tap-major, then channel: the column order of a packed 3x3 weight)."""
    n, h, w, c = x.shape
    _, ho, wo, off_c = offset.shape
    assert offset.dtype == torch.float32 and offset.is_contiguous() and x.is_contiguous()
    assert off_c == 2 * deform_groups * kh * kw
    col = torch.empty(n * ho * wo, kh * kw * c, dtype=x.dtype, device=x.device)
    L.check(L.load().ldb200_deform_im2col(L.ptr(x), L.ptr(offset), L.ptr(col), n, h, w, c, ho, wo, kh, kw, int(stride),
                                          int(pad), int(dilation), int(deform_groups), off_c, _dt(x.dtype), L.stream_ptr()),
            'ldb200_deform_im2col')
    return col


def maxpool3x3s2(x):
    n, h, w, c = x.shape
    out = torch.empty(n, (h - 1) // 2 + 1, (w - 1) // 2 + 1, c, dtype=x.dtype, device=x.device)
    L.check(L.load().ldb200_maxpool3x3s2(L.ptr(x), L.ptr(out), n, h, w, c, _dt(x.dtype), L.stream_ptr()), 'ldb200_maxpool3x3s2')
    return out


def gn_relu_fwd(x, gamma, beta, groups=32, eps=1e-5, relu=True):
    """x [N,H,W,C] bf16 -> (y, stats[N,G,2])."""
    n, h, w, c = x.shape
    lib = L.load()
    stats = torch.empty(n, groups, 2, dtype=torch.float32, device=x.device)
    y = torch.empty_like(x)
    L.check(lib.ldb200_gn_stats(L.ptr(x), L.ptr(stats), n, h * w, c, groups, _dt(x.dtype), L.stream_ptr()), 'ldb200_gn_stats')
    L.check(lib.ldb200_gn_apply(L.ptr(x), L.ptr(stats), L.ptr(gamma), L.ptr(beta), L.ptr(y), n, h * w, c, groups,
                                float(eps), int(relu), _dt(x.dtype), L.stream_ptr()), 'ldb200_gn_apply')
    return y, stats


def gn_apply_multi(xs, stats, gamma, beta, groups=32, eps=1e-5, relu=True):
    """y_l = relu(GroupNorm(x_l)) for all levels in ONE launch; stats[l]: [N, G, 2] (from `gn_stats` or from the
    producing conv's epilogue)."""
    n = len(xs)
    segs = (L.GnSeg * n)()
    ys = []
    c = xs[0].shape[-1]
    for i, x in enumerate(xs):
        assert x.is_contiguous() and x.shape[-1] == c and x.dtype == xs[0].dtype
        y = torch.empty_like(x)
        segs[i].x, segs[i].out, segs[i].stats = x.data_ptr(), y.data_ptr(), stats[i].data_ptr()
        segs[i].N, segs[i].HW = x.shape[0], x.shape[1] * x.shape[2]
        ys.append(y)
    L.check(L.load().ldb200_gn_apply_multi(segs, n, L.ptr(gamma), L.ptr(beta), c, groups, float(eps), int(relu),
                                           _dt(xs[0].dtype), L.stream_ptr()), 'ldb200_gn_apply_multi')
    return ys


def gn_relu_bwd_multi(dys, ys, xs, stats, gamma, dgamma, dbeta, groups=32, eps=1e-5, beta=None):
    """Backward of `gn_apply_multi` (relu=True) for all levels in two launches -> list of dx; dgamma / dbeta [C] f32
    accumulated in place.  ys=None (+ beta): the ReLU mask is recomputed from x instead of being read from y."""
    if ys is None:
        assert beta is not None
        ys = [None] * len(xs)
    n = len(xs)
    segs = (L.GnSeg * n)()
    c = xs[0].shape[-1]
    tot = sum(x.shape[0] for x in xs) * groups * 2
    red = torch.zeros(tot, dtype=torch.float32, device=xs[0].device)     # one fill for all levels
    dxs, o = [], 0
    for i, (dy, y, x) in enumerate(zip(dys, ys, xs)):
        assert dy.is_contiguous() and x.is_contiguous() and dy.dtype == x.dtype
        assert y is None or (y.is_contiguous() and y.dtype == x.dtype)
        dx = torch.empty_like(x)
        k = x.shape[0] * groups * 2
        segs[i].dy, segs[i].x, segs[i].out = dy.data_ptr(), x.data_ptr(), dx.data_ptr()
        segs[i].y = None if y is None else y.data_ptr()
        segs[i].stats, segs[i].red_ws, segs[i].red_zeroed = stats[i].data_ptr(), red[o:o + k].data_ptr(), 1
        segs[i].N, segs[i].HW = x.shape[0], x.shape[1] * x.shape[2]
        o += k
        dxs.append(dx)
    L.check(L.load().ldb200_gn_relu_bwd_multi(segs, n, L.ptr(gamma), L.ptr(beta), L.ptr(dgamma), L.ptr(dbeta), c, groups, float(eps),
                                              _dt(xs[0].dtype), L.stream_ptr()), 'ldb200_gn_relu_bwd_multi')
    del red
    return dxs


def gn_relu_bwd(dy, y, x, stats, gamma, dgamma, dbeta, groups=32, eps=1e-5):
    """-> dx [N,H,W,C] bf16; dgamma/dbeta [C] f32 accumulated in place."""
    n, h, w, c = x.shape
    ws = torch.empty(n, groups, 2, dtype=torch.float32, device=x.device)
    dx = torch.empty_like(x)
    assert dy.dtype == y.dtype == x.dtype
    L.check(L.load().ldb200_gn_relu_bwd(L.ptr(dy), L.ptr(y), L.ptr(x), L.ptr(stats), L.ptr(gamma), L.ptr(ws),
                                        L.ptr(dgamma), L.ptr(dbeta), L.ptr(dx), n, h * w, c, groups, float(eps),
                                        _dt(x.dtype), L.stream_ptr()), 'ldb200_gn_relu_bwd')
    return dx


def upsample_add_bwd(dfine, dcoarse):
    n, hf, wf, c = dfine.shape
    _, hc, wc, _ = dcoarse.shape
    assert dfine.dtype == dcoarse.dtype
    L.check(L.load().ldb200_upsample_add_bwd(L.ptr(dfine), L.ptr(dcoarse), n, hf, wf, hc, wc, c, _dt(dfine.dtype),
                                             L.stream_ptr()), 'ldb200_upsample_add_bwd')
    return dcoarse


def relu_bwd(dy, y, out=None):
    out = torch.empty_like(dy) if out is None else out
    assert dy.dtype == y.dtype
    L.check(L.load().ldb200_relu_bwd(L.ptr(dy), L.ptr(y), L.ptr(out), dy.numel(), _dt(dy.dtype), L.stream_ptr()),
            'ldb200_relu_bwd')
    return out


def colsum(x2d, out):
    """out[c] += sum over rows of x2d [M, C] (bf16 or f32, row stride = x2d.stride(0))."""
    assert x2d.dim() == 2 and x2d.stride(1) == 1 and out.dtype == torch.float32
    L.check(L.load().ldb200_colsum(C.c_void_p(x2d.data_ptr()), L.ptr(out), x2d.shape[0], x2d.shape[1], x2d.stride(0),
                                   _dt(x2d.dtype), L.stream_ptr()), 'ldb200_colsum')
    return out


def sgd_flat(w, g, buf, lr, momentum=0.9, weight_decay=1e-4, grad_scale=1.0, first_step=False):
    L.check(L.load().ldb200_sgd_flat(L.ptr(w), L.ptr(g), L.ptr(buf), w.numel(), float(lr), float(momentum),
                                     float(weight_decay), float(grad_scale), int(first_step), L.stream_ptr()),
            'ldb200_sgd_flat')


def conv2d_wgrad_tc(dy, x, dw, kh=1, kw=1, stride=1, pad=0, scale=None, w_master=None, d_scale=None):
    """dw [Cout, kh*kw*Cin] f32 (packed tap-major layout) += scale[co] * wgrad(dy [N,HO,WO,Cout], x [N,H,W,Cin])
    (bf16 NHWC); optionally d_scale[co] += <w_master[co], wgrad[co]>."""
    assert dy.dtype == x.dtype and dw.dtype == torch.float32 and dw.is_contiguous()
    n, ho, wo, cout = dy.shape
    _, h, w_, cin = x.shape
    d = L.WgradDesc()
    keep = None
    if dy.dtype == torch.float32:     # fp32-accurate mode: both operands as bf16 planes
        keep = (split_bf16(dy), split_bf16(x))
        dy = keep[0][0].view(n, ho, wo, -1)[..., :cout]
        x = keep[1][0].view(n, h, w_, -1)[..., :cin]
        d.split = keep[0].shape[0]
    assert dy.dtype == torch.bfloat16
    d.dy, d.N, d.HO, d.WO, d.Cout, d.dy_ld = dy.data_ptr(), n, ho, wo, cout, dy.stride(2)
    d.x, d.H, d.W, d.Cin, d.x_ld = x.data_ptr(), h, w_, cin, x.stride(2)
    taps = [(s - pad, r - pad, (r * kw + s) * cin) for r in range(kh) for s in range(kw)]
    d.ntaps = len(taps)
    for i, (dx, dy_, wk) in enumerate(taps):
        d.tap_dx[i], d.tap_dy[i], d.tap_wk[i] = dx, dy_, wk
    d.sx = d.sy = stride
    d.dw, d.w_k = dw.data_ptr(), dw.shape[1]
    if scale is not None:
        assert scale.dtype == torch.float32 and scale.is_contiguous()
        d.scale = scale.data_ptr()
    if d_scale is not None:
        assert w_master is not None and w_master.dtype == torch.float32 and w_master.shape == dw.shape \
            and w_master.is_contiguous() and d_scale.dtype == torch.float32
        d.w_master, d.d_scale = w_master.data_ptr(), d_scale.data_ptr()
    if d.split and PARITY_TWO_PASS:      # leading plane product, then the corrections: two accumulating launches
        for sp in (1, 2):
            d.split_pass = sp
            L.check(L.load().ldb200_conv2d_wgrad_tc(C.byref(d), L.stream_ptr()), 'ldb200_conv2d_wgrad_tc')
        return dw
    L.check(L.load().ldb200_conv2d_wgrad_tc(C.byref(d), L.stream_ptr()), 'ldb200_conv2d_wgrad_tc')
    return dw


def unpack_conv_weight_grad(dw_packed, cout, cin, kh, kw):
    """[Cout, kh*kw*Cin] -> [Cout, Cin, kh, kw] (PyTorch OIHW view of the packed gradient)."""
    return dw_packed.view(cout, kh, kw, cin).permute(0, 3, 1, 2)


def sgd_flat_dev(w, g, buf, hyper_dev):
    """Graph-capturable SGD: hyper_dev = f32[5] device {lr, momentum, wd, grad_scale, first_step}."""
    L.check(L.load().ldb200_sgd_flat_dev(L.ptr(w), L.ptr(g), L.ptr(buf), w.numel(), L.ptr(hyper_dev), L.stream_ptr()),
            'ldb200_sgd_flat_dev')

"""CPU oracle: an independent fp32 restatement (torch-CPU tensor arithmetic) of the reference's
LD train-step math.  TEST INFRASTRUCTURE ONLY -- only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s cpu_baseline / `--impl reference` legs may import this; it is the checker, never the
product (`ld_b200/` fails loudly if its CUDA library is missing).

Pinned against the UNMODIFIED reference run through `oracle/mmcv_standin.py` (see
`oracle/make_golden.py` -> `tests/golden/*.pt` and `tests/test_oracle_vs_reference.py`): the
reference's own tests hold no golden vectors for this path, so the goldens are
outputs of the reference code itself on seeded inputs.

Each function cites the reference file:line it restates (paths relative to /root/reference).
Target generation is written dense and sort-free (boolean candidate masks instead of the
reference's scatter/sort) -- SURVEY.md A.3-A.5 -- so it is an independent derivation, not a port.
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F

INF = 100000000.0


# =========================================================================== anchors
def level_anchors(h, w, stride, device='cpu'):
    """mmdet/core/anchor/anchor_generator.py:142-185,233-270 with octave_base_scale=8,
    ratios=[1.0], scales_per_octave=1, center_offset=0: one square anchor of side 8*stride per
    cell, centred at (ix*s, iy*s), x fastest."""
    half = 4.0 * stride
    xs = torch.arange(w, device=device, dtype=torch.float32) * stride
    ys = torch.arange(h, device=device, dtype=torch.float32) * stride
    cx = xs.repeat(h)
    cy = ys.repeat_interleave(w)
    return torch.stack([cx - half, cy - half, cx + half, cy + half], 1)


def level_valid_flags(h, w, stride, pad_h, pad_w, device='cpu'):
    """anchor_generator.py:272-328: cell valid iff inside ceil(pad/stride)."""
    vh = min(int(math.ceil(pad_h / stride)), h)
    vw = min(int(math.ceil(pad_w / stride)), w)
    fx = torch.arange(w, device=device) < vw
    fy = torch.arange(h, device=device) < vh
    return (fy[:, None] & fx[None, :]).reshape(-1)


# =========================================================================== box geometry
def pair_overlaps(a, b, mode='iou', eps=1e-6):
    """mmdet/core/bbox/iou_calculators/iou2d_calculator.py:43-188, pairwise [m,n].
    NB: for mode 'diou' the reference's base term is IoF = overlap/area(a)
    (the `else: union = area1` branch :146-149), not IoU."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, None, :2], b[None, :, :2])
    rb = torch.min(a[:, None, 2:], b[None, :, 2:])
    wh = (rb - lt).clamp(min=0)
    ov = wh[..., 0] * wh[..., 1]
    if mode in ('iou', 'giou'):
        union = area_a[:, None] + area_b[None, :] - ov
    else:
        union = area_a[:, None].expand_as(ov)
    union = torch.max(union, union.new_tensor(eps))
    base = ov / union
    if mode in ('iou', 'iof'):
        return base
    elt = torch.min(a[:, None, :2], b[None, :, :2])
    erb = torch.max(a[:, None, 2:], b[None, :, 2:])
    ewh = (erb - elt).clamp(min=0)
    if mode == 'giou':
        ea = torch.max(ewh[..., 0] * ewh[..., 1], union.new_tensor(eps))
        return base - (ea - union) / ea
    # 'diou'
    left = (b[None, :, 0] + b[None, :, 2] - (a[:, None, 0] + a[:, None, 2])) ** 2 / 4
    right = (b[None, :, 1] + b[None, :, 3] - (a[:, None, 1] + a[:, None, 3])) ** 2 / 4
    rho2 = left + right
    c2 = torch.max(ewh[..., 0] ** 2 + ewh[..., 1] ** 2, union.new_tensor(eps))
    return base - rho2 / c2


def aligned_overlaps(a, b, mode='iou', eps=1e-6):
    """iou2d_calculator.py:111-143 (is_aligned=True) for 'iou' / 'giou'."""
    area_a = (a[:, 2] - a[:, 0]) * (a[:, 3] - a[:, 1])
    area_b = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = torch.max(a[:, :2], b[:, :2])
    rb = torch.min(a[:, 2:], b[:, 2:])
    wh = (rb - lt).clamp(min=0)
    ov = wh[:, 0] * wh[:, 1]
    union = torch.max(area_a + area_b - ov, ov.new_tensor(eps))
    iou = ov / union
    if mode == 'iou':
        return iou
    elt = torch.min(a[:, :2], b[:, :2])
    erb = torch.max(a[:, 2:], b[:, 2:])
    ewh = (erb - elt).clamp(min=0)
    ea = torch.max(ewh[:, 0] * ewh[:, 1], ov.new_tensor(eps))
    return iou - (ea - union) / ea


# =========================================================================== targets
def atss_candidates(anchors, num_level, gts, topk=9):
    """Boolean [A,G] mask of the per-level top-k closest anchors for each gt
    (mmdet/core/bbox/assigners/atss_assigner.py:94-127); ties -> lowest index (torch CPU topk)."""
    acx = (anchors[:, 0] + anchors[:, 2]) / 2.0
    acy = (anchors[:, 1] + anchors[:, 3]) / 2.0
    gcx = (gts[:, 0] + gts[:, 2]) / 2.0
    gcy = (gts[:, 1] + gts[:, 3]) / 2.0
    d = ((acx[:, None] - gcx[None, :]).pow(2) + (acy[:, None] - gcy[None, :]).pow(2)).sqrt()
    cand = torch.zeros_like(d, dtype=torch.bool)
    cand_idx = []
    s = 0
    for n in num_level:
        k = min(topk, n)
        if k > 0:
            # stable ascending sort == "k smallest, lowest index first on ties"
            order = torch.sort(d[s:s + n], dim=0, stable=True)[1][:k] + s
            cand.scatter_(0, order, True)
            cand_idx.append(order)
        s += n
    return cand, torch.cat(cand_idx, 0), acx, acy


def atss_threshold(iou, cand_idx):
    """mean + unbiased std of the candidate IoUs per gt (atss_assigner.py:131-134)."""
    g = iou.shape[1]
    c = iou[cand_idx, torch.arange(g)]
    return c.mean(0) + c.std(0)


def ld_targets_single(anchors, valid, num_level, gts, gt_labels, num_classes=80, topk=9,
                      imitation='finegrained'):
    """One image: `LDHead._get_target_single` (ld_head.py:449-577) = ATSS assign
    (atss_assigner.py:33-181) + PseudoSampler + `get_vlr_region` (:183-298) +
    `get_im_region` (ld_head.py:580-611) + unmap.  allowed_border=-1 => inside == valid.

    Returns dict(labels i64 [A], label_weights f32, bbox_targets f32 [A,4], vlr f32, im f32,
                 gt_inds i64 [A] (0 = bg, k+1 = gt k), num_pos int)."""
    A = anchors.shape[0]
    idx = valid.nonzero().squeeze(1)
    a = anchors[idx]
    nl = []
    s = 0
    for n in num_level:
        nl.append(int(valid[s:s + n].sum()))
        s += n
    G = gts.shape[0]
    assert G >= 1, 'reference get_vlr_region returns a non-tensor for G==0 (atss_assigner.py:203-216)'
    iou = pair_overlaps(a, gts, 'iou')
    cand, cand_idx, acx, acy = atss_candidates(a, nl, gts, topk)
    thr = atss_threshold(iou, cand_idx)
    # --- positives (A.3)
    l_ = acx[:, None] - gts[None, :, 0]
    t_ = acy[:, None] - gts[None, :, 1]
    r_ = gts[None, :, 2] - acx[:, None]
    b_ = gts[None, :, 3] - acy[:, None]
    inside = torch.stack([l_, t_, r_, b_], 0).min(0)[0] > 0.01
    pos = cand & (iou >= thr[None, :]) & inside
    masked = torch.where(pos, iou, iou.new_full((), -INF))
    mx, am = masked.max(1)
    gt_inds = torch.where(mx != -INF, am + 1, torch.zeros_like(am))
    # --- VLR (A.4): band on the IoF-based "diou", over ALL anchors
    diou = pair_overlaps(a, gts, 'diou')
    band = (diou < thr[None, :]) & (diou >= 0.25 * thr[None, :])
    vm = torch.where(band, iou, iou.new_full((), -INF)).max(1)[0]
    vlr = torch.where(vm != -INF, vm, torch.zeros_like(vm))
    # --- imitation region (A.5)
    if imitation == 'finegrained':
        im = (iou > 0.5 * iou.max(0)[0][None, :]).any(1).float()
    else:
        im = ((acx[:, None] > gts[None, :, 0]) & (acx[:, None] < gts[None, :, 2]) &
              (acy[:, None] > gts[None, :, 1]) & (acy[:, None] < gts[None, :, 3])).any(1).float()
    is_pos = gt_inds > 0
    lab = torch.where(is_pos, gt_labels[(gt_inds - 1).clamp(min=0)],
                      torch.full_like(gt_inds, num_classes))
    bt = torch.where(is_pos[:, None], gts[(gt_inds - 1).clamp(min=0)], torch.zeros_like(a))
    out = dict(labels=torch.full((A,), num_classes, dtype=torch.long),
               label_weights=torch.zeros(A), bbox_targets=torch.zeros(A, 4),
               vlr=torch.zeros(A), im=torch.zeros(A), gt_inds=torch.zeros(A, dtype=torch.long))
    out['labels'][idx] = lab
    out['label_weights'][idx] = 1.0
    out['bbox_targets'][idx] = bt
    out['vlr'][idx] = vlr
    out['im'][idx] = im
    out['gt_inds'][idx] = gt_inds
    out['num_pos'] = int(is_pos.sum())
    return out


def ld_targets(featmap_sizes, strides, img_metas, gt_bboxes, gt_labels, **kw):
    """Batch version (ld_head.py:377-447): per-level stacked tensors [N, A_l(,4)] + num_total_pos."""
    anchors = [level_anchors(h, w, s) for (h, w), s in zip(featmap_sizes, strides)]
    num_level = [a.shape[0] for a in anchors]
    flat = torch.cat(anchors, 0)
    per_img = []
    for meta, gb, gl in zip(img_metas, gt_bboxes, gt_labels):
        ph, pw = meta['pad_shape'][:2]
        valid = torch.cat([level_valid_flags(h, w, s, ph, pw)
                           for (h, w), s in zip(featmap_sizes, strides)])
        per_img.append(ld_targets_single(flat, valid, num_level, gb, gl, **kw))
    num_total_pos = sum(max(t['num_pos'], 1) for t in per_img)
    out = {}
    for key in ('labels', 'label_weights', 'bbox_targets', 'vlr', 'im', 'gt_inds'):
        stacked = torch.stack([t[key] for t in per_img], 0)
        out[key] = list(stacked.split(num_level, dim=1))
    out['anchors'] = anchors
    out['num_total_pos'] = num_total_pos
    return out


# =========================================================================== losses
def qfl_rows(logits, labels, score, beta=2.0):
    """quality_focal_loss (mmdet/models/losses/gfocal_loss.py:34-49), per-anchor row sums."""
    sig = logits.sigmoid()
    loss = F.binary_cross_entropy_with_logits(logits, torch.zeros_like(logits),
                                              reduction='none') * sig.pow(beta)
    C = logits.shape[1]
    pos = ((labels >= 0) & (labels < C)).nonzero().squeeze(1)
    if pos.numel():
        pl = labels[pos]
        x = logits[pos, pl]
        sf = score[pos] - sig[pos, pl]
        lp = F.binary_cross_entropy_with_logits(x, score[pos], reduction='none') * sf.abs().pow(beta)
        loss = loss.index_put((pos, pl), lp)
    return loss.sum(1)


def qfl_prob_rows(prob, labels, score, beta=2.0):
    """quality_focal_loss with use_sigmoid=False (gfocal_loss.py:29-49; the LDv2 / GFLv2 configs): `prob` is already
    the joint class-quality score sigmoid(cls_feat) * quality (gfocal_head.py:215), supervised by plain
    F.binary_cross_entropy.  C = num_classes + 1 there (anchor_head.py use_sigmoid_cls=False), so background rows
    (label == num_classes) take the "positive" branch with score 0 -- numerically the zero-label formula."""
    loss = F.binary_cross_entropy(prob, torch.zeros_like(prob), reduction='none') * prob.pow(beta)
    C = prob.shape[1]
    pos = ((labels >= 0) & (labels < C)).nonzero().squeeze(1)
    if pos.numel():
        pl = labels[pos]
        x = prob[pos, pl]
        sf = score[pos] - x
        lp = F.binary_cross_entropy(x, score[pos], reduction='none') * sf.abs().pow(beta)
        loss = loss.index_put((pos, pl), lp)
    return loss.sum(1)


def dfl_rows(logits, target):
    """distribution_focal_loss (gfocal_loss.py:68-74)."""
    dl = target.long()
    dr = dl + 1
    wl = dr.float() - target
    wr = target - dl.float()
    return F.cross_entropy(logits, dl, reduction='none') * wl + \
        F.cross_entropy(logits, dr, reduction='none') * wr


def kd_kl_rows(student, teacher, T):
    """knowledge_distillation_kl_div_loss (mmdet/models/losses/kd_loss.py:28-34)."""
    p_t = F.softmax(teacher / T, dim=1).detach()
    return F.kl_div(F.log_softmax(student / T, dim=1), p_t, reduction='none').mean(1) * (T * T)


def integral(reg, reg_max=16):
    """Integral (mmdet/models/dense_heads/gfl_head.py:32-44)."""
    p = F.softmax(reg.reshape(-1, reg_max + 1), dim=1)
    proj = torch.linspace(0, reg_max, reg_max + 1, dtype=reg.dtype, device=reg.device)
    return F.linear(p, proj).reshape(-1, 4)


def nms_topk(boxes, scores, iou_thr, k):
    """The first `k` indices `torchvision.ops.nms(boxes, scores, iou_thr)` keeps (ld_head.py:637
    `torch.ops.torchvision.nms(gibox, giscore, 0.3)[:10]`).  torchvision 0.x CPU kernel
    (torchvision/csrc/ops/cpu/nms_kernel.cpp, not vendored in /root/reference): candidates in
    stable descending score order; a kept box i suppresses every later j with
    inter / (area_i + area_j - inter) > iou_thr, areas = (x2-x1)*(y2-y1), inter from clamped
    widths/heights, all in the dtype of `boxes`."""
    order = torch.sort(scores, descending=True, stable=True)[1]
    x1, y1, x2, y2 = boxes.unbind(1)
    areas = (x2 - x1) * (y2 - y1)
    suppressed = torch.zeros(boxes.shape[0], dtype=torch.bool)
    keep = []
    pos = 0
    n = boxes.shape[0]
    while len(keep) < k and pos < n:
        # next unsuppressed candidate in score order
        rest = order[pos:]
        alive = (~suppressed[rest]).nonzero()
        if alive.numel() == 0:
            break
        pos += int(alive[0])
        i = int(order[pos])
        pos += 1
        keep.append(i)
        w = (torch.minimum(x2[i], x2) - torch.maximum(x1[i], x1)).clamp(min=0)
        h = (torch.minimum(y2[i], y2) - torch.maximum(y1[i], y1)).clamp(min=0)
        inter = w * h
        ovr = inter / (areas[i] + areas - inter)
        suppressed |= ovr > iou_thr
        suppressed[i] = True
    return torch.tensor(keep, dtype=torch.long)


def gi_region(t_cls, cls, anchors, reg, t_reg, stride, reg_max=16, iou_thr=0.3, k=10, raw_scores=False):
    """`LDHead.get_gi_region` (ld_head.py:613-638): General-Instance boxes of one level.  Per location the GI
    score is max_c |sigmoid(teacher) - sigmoid(student)|, the GI box the teacher's decoded box where the teacher
    scores higher at that class, else the student's; NMS (IoU 0.3) over ALL locations of the level (the images of
    the batch share one coordinate frame here, as in the reference) and the first 10 keeps are returned."""
    if raw_scores:   # LDv2Head.get_gi_region (ld_gflv2.py): no sigmoid -- teacher LOGITS against student joint scores
        ts, ss = t_cls.detach(), cls.detach()
    else:
        ts, ss = t_cls.detach().sigmoid(), cls.detach().sigmoid()
    ctr = torch.stack([(anchors[:, 0] + anchors[:, 2]) / 2, (anchors[:, 1] + anchors[:, 3]) / 2], 1) / stride
    sd, td = integral(reg.detach(), reg_max), integral(t_reg.detach(), reg_max)

    def d2b(d):   # distance2bbox (core/bbox/transforms.py:119-156), no max_shape
        return torch.stack([ctr[:, 0] - d[:, 0], ctr[:, 1] - d[:, 1], ctr[:, 0] + d[:, 2], ctr[:, 1] + d[:, 3]], 1)

    z = ts - ss
    giscore, index = z.abs().max(1)
    teacher_bigger = z.gather(1, index[:, None]).squeeze(1) >= 0
    gibox = torch.where(teacher_bigger[:, None], d2b(td), d2b(sd))
    return nms_topk(gibox, giscore, iou_thr, k)


DEFAULT_LOSS_CFG = dict(w_qfl=1.0, beta=2.0, w_giou=2.0, w_dfl=0.25, w_ld=0.25, T_ld=10.0,
                        w_ld_vlr=0.25, T_ld_vlr=10.0, w_kd=10.0, T_kd=2.0, w_im=2.0, reg_max=16)


def ld_loss_level(anchors, cls, reg, labels, label_w, bbox_t, stride, t_reg, t_cls, x, t_x,
                  vlr, im, num_total_samples, cfg=DEFAULT_LOSS_CFG, imitation='finegrained', gi_out=None,
                  cls_feat=None, num_classes=None):
    """`LDHead.loss_single` (ld_head.py:116-282) on flattened NHWC rows.

    anchors [R,4]; cls [R,C]; reg/t_reg [R,4*(reg_max+1)]; t_cls [R,C]; x,t_x [R,F].
    Returns the 8 (pre-avg_factor) loss terms and sum of weight_targets."""
    # LDv2Head.loss_single (ld_gflv2.py:116-285) when `cls_feat` is given: `cls` is then the joint score
    # sigmoid(cls_feat) * quality (a probability), `t_cls` the teacher's cls_feat LOGITS (ld_gflv2.py:327)
    v2 = cls_feat is not None
    C = cls.shape[1] if num_classes is None else num_classes    # bg_class_ind = self.num_classes (ld_head.py:163)
    rm = cfg['reg_max']
    pos = ((labels >= 0) & (labels < C)).nonzero().squeeze(1)
    rem = (vlr > 0).nonzero().squeeze(1)
    fg = (im > 0).nonzero().squeeze(1)
    zero = reg.sum() * 0
    score = label_w.new_zeros(labels.shape)
    if imitation == 'gibox':   # ld_head.py:170-176
        fg = gi_region(t_cls, cls, anchors, reg, t_reg, stride, rm, raw_scores=v2)
        if gi_out is not None:
            gi_out.append(fg)
    loss_im = cfg['w_im'] * F.mse_loss(x[fg], t_x[fg]) if fg.numel() else zero
    if pos.numel():
        a = anchors[pos]
        ctr = torch.stack([(a[:, 0] + a[:, 2]) / 2, (a[:, 1] + a[:, 3]) / 2], 1) / stride
        wt = (cls.detach() if v2 else cls.detach().sigmoid()).max(1)[0][pos]   # ld_gflv2.py:200 / ld_head.py:198
        d = integral(reg[pos], rm)
        pred = torch.stack([ctr[:, 0] - d[:, 0], ctr[:, 1] - d[:, 1],
                            ctr[:, 0] + d[:, 2], ctr[:, 1] + d[:, 3]], 1)
        tgt = bbox_t[pos] / stride
        score[pos] = aligned_overlaps(pred.detach(), tgt, 'iou')
        corners = torch.stack([ctr[:, 0] - tgt[:, 0], ctr[:, 1] - tgt[:, 1],
                               tgt[:, 2] - ctr[:, 0], tgt[:, 3] - ctr[:, 1]], 1
                              ).clamp(min=0, max=rm - 0.1).reshape(-1)
        w4 = wt[:, None].expand(-1, 4).reshape(-1)
        pc = reg[pos].reshape(-1, rm + 1)
        sc = t_reg[pos].reshape(-1, rm + 1)
        loss_bbox = cfg['w_giou'] * ((1 - aligned_overlaps(pred, tgt, 'giou')) * wt).sum() / 1.0
        loss_dfl = cfg['w_dfl'] * (dfl_rows(pc, corners) * w4).sum() / 4.0
        loss_ld = cfg['w_ld'] * (kd_kl_rows(pc, sc, cfg['T_ld']) * w4).sum() / 4.0
        loss_kd = cfg['w_kd'] * (kd_kl_rows((cls_feat if v2 else cls)[pos], t_cls[pos], cfg['T_kd']) * label_w[pos]
                                 ).sum() / pos.numel()
        wsum = wt.sum()
    else:
        loss_bbox = loss_dfl = loss_ld = loss_kd = zero
        loss_im = zero  # ld_head.py:251 zeroes loss_im too when the level has no positive
        wsum = reg.new_tensor(0.)
    if rem.numel():
        rw = vlr[rem][:, None].expand(-1, 4).reshape(-1)
        loss_ld_vlr = cfg['w_ld_vlr'] * (kd_kl_rows(reg[rem].reshape(-1, rm + 1),
                                                    t_reg[rem].reshape(-1, rm + 1),
                                                    cfg['T_ld_vlr']) * rw).sum() / 16.0
        loss_kd_neg = 0 * (cfg['w_kd'] * (kd_kl_rows(cls[rem], t_cls[rem], cfg['T_kd'])
                                          * label_w[rem]).sum() / rem.numel())
    else:
        loss_ld_vlr = loss_kd_neg = zero
    qfl = qfl_prob_rows if v2 else qfl_rows
    loss_cls = cfg['w_qfl'] * (qfl(cls, labels, score, cfg['beta']) * label_w).sum() / num_total_samples
    return dict(loss_cls=loss_cls, loss_bbox=loss_bbox, loss_dfl=loss_dfl, loss_ld=loss_ld,
                loss_ld_vlr=loss_ld_vlr, loss_kd=loss_kd, loss_kd_neg=loss_kd_neg,
                loss_im=loss_im), wsum


def nhwc_rows(t):
    """[N,C,H,W] -> [N*H*W, C] (ld_head.py:144-154)."""
    return t.permute(0, 2, 3, 1).reshape(-1, t.shape[1])


def ld_loss(cls_scores, bbox_preds, t_cls_scores, t_bbox_preds, feats, t_feats, targets,
            strides, cfg=DEFAULT_LOSS_CFG, world_size=1, imitation='finegrained', gi_out=None, cls_feats=None,
            num_classes=None):
    """`LDHead.loss` (ld_head.py:284-375) given NCHW per-level tensors and `ld_targets` output.
    Single process: reduce_mean is the identity."""
    nts = max(float(targets["num_total_pos"]), 1.0)
    per = []
    wsum = 0
    n = cls_scores[0].shape[0]
    for l, s in enumerate(strides):
        anc = targets['anchors'][l].repeat(n, 1)
        d, w = ld_loss_level(anc, nhwc_rows(cls_scores[l]), nhwc_rows(bbox_preds[l]),
                             targets['labels'][l].reshape(-1),
                             targets['label_weights'][l].reshape(-1),
                             targets['bbox_targets'][l].reshape(-1, 4), float(s),
                             nhwc_rows(t_bbox_preds[l]), nhwc_rows(t_cls_scores[l]),
                             nhwc_rows(feats[l]), nhwc_rows(t_feats[l]),
                             targets['vlr'][l].reshape(-1), targets['im'][l].reshape(-1),
                             nts, cfg, imitation, gi_out,
                             None if cls_feats is None else nhwc_rows(cls_feats[l]), num_classes)
        per.append(d)
        wsum = wsum + w
    avg = float(wsum + 1e-6)
    out = {k: [p[k] for p in per] for k in per[0]}
    out['loss_bbox'] = [v / avg for v in out['loss_bbox']]
    out['loss_dfl'] = [v / avg for v in out['loss_dfl']]
    return out


def parse_losses(losses):
    """`BaseDetector._parse_losses` (mmdet/models/detectors/base.py:185-218), single process."""
    log = {k: sum(v.mean() for v in vs) for k, vs in losses.items()}
    total = sum(v for k, v in log.items() if 'loss' in k)
    log['loss'] = total
    return total, {k: float(v.detach()) for k, v in log.items()}


# =========================================================================== inference
def get_bboxes(cls_scores, bbox_preds, strides, img_shapes, scale_factors=None, nms_pre=1000, score_thr=0.05,
               iou_thr=0.6, max_per_img=100, reg_max=16):
    """`GFLHead._get_bboxes` (gfl_head.py:354-451) + `multiclass_nms` type 'nms' (bbox_nms.py:70-193): per-level top
    nms_pre anchors by max class score, Integral * stride, distance2bbox clipped to the image, score threshold, per-class
    NMS through mmcv's coordinate-offset trick, top max_per_img.  -> list of (dets [k,5], labels [k]) per image."""
    from .mmcv_standin import batched_nms
    n = cls_scores[0].shape[0]
    out = []
    for i in range(n):
        boxes_l, scores_l = [], []
        for cls, reg, st in zip(cls_scores, bbox_preds, strides):
            h, w = cls.shape[2:]
            sc = cls[i].permute(1, 2, 0).reshape(h * w, -1).sigmoid()
            d = integral(reg[i].permute(1, 2, 0).reshape(h * w, -1), reg_max) * st
            anchors = level_anchors(h, w, st)
            if nms_pre > 0 and sc.shape[0] > nms_pre:
                top = sc.max(-1)[0].topk(nms_pre)[1]
                sc, d, anchors = sc[top], d[top], anchors[top]
            cx, cy = (anchors[:, 0] + anchors[:, 2]) / 2, (anchors[:, 1] + anchors[:, 3]) / 2
            b = torch.stack([cx - d[:, 0], cy - d[:, 1], cx + d[:, 2], cy + d[:, 3]], -1)
            ih, iw = img_shapes[i][:2]
            b = torch.stack([b[:, 0].clamp(0, iw), b[:, 1].clamp(0, ih), b[:, 2].clamp(0, iw), b[:, 3].clamp(0, ih)], -1)
            boxes_l.append(b)
            scores_l.append(sc)
        boxes, scores = torch.cat(boxes_l), torch.cat(scores_l)
        if scale_factors is not None:
            boxes = boxes / boxes.new_tensor(scale_factors[i])
        C = scores.shape[1]
        bb = boxes[:, None].expand(-1, C, 4).reshape(-1, 4)
        ss = scores.reshape(-1)
        ll = torch.arange(C).view(1, -1).expand_as(scores).reshape(-1)
        valid = (ss > score_thr).nonzero(as_tuple=False).squeeze(1)
        bb, ss, ll = bb[valid], ss[valid], ll[valid]
        if bb.numel() == 0:
            out.append((boxes.new_zeros(0, 5), ll.new_zeros(0)))
            continue
        dets, keep = batched_nms(bb, ss, ll, dict(type='nms', iou_threshold=iou_thr))
        out.append((dets[:max_per_img], ll[keep][:max_per_img]))
    return out


# =========================================================================== network port
class _Bottleneck(nn.Module):
    """mmdet/models/backbones/resnet.py:95-300, style='pytorch' (stride on the 3x3)."""
    expansion = 4

    def __init__(self, inp, planes, stride, downsample, groups=1, base_width=4, dcn=False):
        """groups / base_width: ResNeXt bottleneck (resnext.py:11-106); dcn: conv2 is a DeformConv2dPack
        (resnet.py:172-184; mmcv semantics restated in oracle/mmcv_standin.py)."""
        super().__init__()
        import math
        width = planes if groups == 1 else math.floor(planes * (base_width / 64)) * groups
        self.conv1 = nn.Conv2d(inp, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        if dcn:
            from .mmcv_standin import DeformConv2dPack
            self.conv2 = DeformConv2dPack(width, width, 3, stride, 1, groups=groups, deform_groups=1)
        else:
            self.conv2 = nn.Conv2d(width, width, 3, stride, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = nn.Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        o = F.relu(self.bn1(self.conv1(x)))
        o = F.relu(self.bn2(self.conv2(o)))
        o = self.bn3(self.conv3(o))
        return F.relu(o + idt)


class _BasicBlock(nn.Module):
    """resnet.py:13-92."""
    expansion = 1

    def __init__(self, inp, planes, stride, downsample):
        super().__init__()
        self.conv1 = nn.Conv2d(inp, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = nn.Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample

    def forward(self, x):
        idt = x if self.downsample is None else self.downsample(x)
        o = F.relu(self.bn1(self.conv1(x)))
        o = self.bn2(self.conv2(o))
        return F.relu(o + idt)


_ARCH = {18: (_BasicBlock, (2, 2, 2, 2)), 34: (_BasicBlock, (3, 4, 6, 3)),
         50: (_Bottleneck, (3, 4, 6, 3)), 101: (_Bottleneck, (3, 4, 23, 3))}


class OracleResNet(nn.Module):
    """resnet.py:303-649 with frozen_stages=1, norm_eval=True (BN always eval)."""

    def __init__(self, depth, groups=1, base_width=4, stage_with_dcn=(False, False, False, False)):
        """groups > 1: ResNeXt (resnext.py:109-153); stage_with_dcn: DCN 3x3 convs in those stages."""
        super().__init__()
        block, layers = _ARCH[depth]
        self.conv1 = nn.Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inp = 64
        for i, n in enumerate(layers):
            planes = 64 * 2 ** i
            stride = 1 if i == 0 else 2
            blocks = []
            for j in range(n):
                st = stride if j == 0 else 1
                ds = None
                if j == 0 and (st != 1 or inp != planes * block.expansion):
                    ds = nn.Sequential(nn.Conv2d(inp, planes * block.expansion, 1, st, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
                kw = dict(groups=groups, base_width=base_width, dcn=stage_with_dcn[i]) if block is _Bottleneck else {}
                blocks.append(block(inp, planes, st, ds, **kw))
                inp = planes * block.expansion
            setattr(self, f'layer{i + 1}', nn.Sequential(*blocks))
        self.out_channels = [64 * 2 ** i * block.expansion for i in range(4)]
        for m in [self.conv1, self.bn1, self.layer1]:
            for p in m.parameters():
                p.requires_grad = False

    def train(self, mode=True):
        super().train(mode)
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()
        return self

    def forward(self, x):
        x = F.max_pool2d(F.relu(self.bn1(self.conv1(x))), 3, 2, 1)
        outs = []
        for i in range(4):
            x = getattr(self, f'layer{i + 1}')(x)
            outs.append(x)
        return outs


class _ConvOnly(nn.Module):
    def __init__(self, i, o, k, s=1, p=0, bias=True):
        super().__init__()
        self.conv = nn.Conv2d(i, o, k, s, p, bias=bias)

    def forward(self, x):
        return self.conv(x)


class OracleFPN(nn.Module):
    """mmdet/models/necks/fpn.py:66-221 with start_level=1, add_extra_convs='on_output', 5 outs."""

    def __init__(self, in_channels, out_channels=256):
        super().__init__()
        self.lateral_convs = nn.ModuleList(_ConvOnly(c, out_channels, 1) for c in in_channels[1:])
        self.fpn_convs = nn.ModuleList(
            [_ConvOnly(out_channels, out_channels, 3, 1, 1) for _ in range(3)] +
            [_ConvOnly(out_channels, out_channels, 3, 2, 1) for _ in range(2)])

    def forward(self, feats):
        lat = [c(f) for c, f in zip(self.lateral_convs, feats[1:])]
        for i in (2, 1):
            lat[i - 1] = lat[i - 1] + F.interpolate(lat[i], size=lat[i - 1].shape[2:],
                                                    mode='nearest')
        outs = [self.fpn_convs[i](lat[i]) for i in range(3)]
        outs.append(self.fpn_convs[3](outs[-1]))
        outs.append(self.fpn_convs[4](outs[-1]))
        return outs


class _ConvGN(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.conv = nn.Conv2d(c, c, 3, 1, 1, bias=False)
        self.gn = nn.GroupNorm(32, c)

    def forward(self, x):
        return F.relu(self.gn(self.conv(x)))


class _Scale(nn.Module):
    def __init__(self):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(1.0))

    def forward(self, x):
        return x * self.scale


class OracleGFLHead(nn.Module):
    """gfl_head.py:102-183."""

    def __init__(self, num_classes=80, c=256, reg_max=16, stacked=4, levels=5):
        super().__init__()
        self.cls_convs = nn.ModuleList(_ConvGN(c) for _ in range(stacked))
        self.reg_convs = nn.ModuleList(_ConvGN(c) for _ in range(stacked))
        self.gfl_cls = nn.Conv2d(c, num_classes, 3, padding=1)
        self.gfl_reg = nn.Conv2d(c, 4 * (reg_max + 1), 3, padding=1)
        self.scales = nn.ModuleList(_Scale() for _ in range(levels))
        self.integral = nn.Module()  # state_dict key parity: bbox_head.integral.project (gfl_head.py:28-30)
        self.integral.register_buffer('project', torch.linspace(0, reg_max, reg_max + 1))

    def forward(self, feats):
        cls_scores, bbox_preds = [], []
        for x, sc in zip(feats, self.scales):
            cf = rf = x
            for m in self.cls_convs:
                cf = m(cf)
            for m in self.reg_convs:
                rf = m(rf)
            cls_scores.append(self.gfl_cls(cf))
            bbox_preds.append(sc(self.gfl_reg(rf)).float())
        return cls_scores, bbox_preds


class OracleGFocalHead(OracleGFLHead):
    """GFLv2 head (gfocal_head.py:112-217): GFLHead + the Distribution-Guided Quality Predictor -- top-`reg_topk`
    probabilities (+ their mean) of each side's bin distribution -> 1x1 conv -> ReLU -> 1x1 conv -> sigmoid = quality,
    cls_score = sigmoid(cls_feat) * quality.  `use_sigmoid=False` in loss_cls makes cls_out_channels = classes + 1."""

    def __init__(self, num_classes=80, c=256, reg_max=16, stacked=4, levels=5, reg_topk=4, reg_channels=64,
                 add_mean=True):
        super().__init__(num_classes + 1, c, reg_max, stacked, levels)
        self.reg_max, self.reg_topk, self.add_mean = reg_max, reg_topk, add_mean
        total_dim = reg_topk + (1 if add_mean else 0)
        self.reg_conf = nn.Sequential(nn.Conv2d(4 * total_dim, reg_channels, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(reg_channels, 1, 1), nn.Sigmoid())

    def forward(self, feats):
        cls_scores, bbox_preds, cls_feats = [], [], []
        for x, sc in zip(feats, self.scales):
            cf = rf = x
            for m in self.cls_convs:
                cf = m(cf)
            for m in self.reg_convs:
                rf = m(rf)
            bp = sc(self.gfl_reg(rf)).float()
            n, _, h, w = bp.shape
            prob = F.softmax(bp.reshape(n, 4, self.reg_max + 1, h, w), dim=2)
            stat = prob.topk(self.reg_topk, dim=2)[0]
            if self.add_mean:
                stat = torch.cat([stat, stat.mean(dim=2, keepdim=True)], dim=2)
            logits = self.gfl_cls(cf)
            cls_scores.append(logits.sigmoid() * self.reg_conf(stat.reshape(n, -1, h, w)))
            bbox_preds.append(bp)
            cls_feats.append(logits)
        return cls_scores, bbox_preds, cls_feats


class OracleGFL(nn.Module):
    def __init__(self, depth, v2=False):
        super().__init__()
        self.backbone = OracleResNet(depth)
        self.neck = OracleFPN(self.backbone.out_channels)
        self.bbox_head = OracleGFocalHead() if v2 else OracleGFLHead()

    def extract_feat(self, img):
        return self.neck(self.backbone(img))


class OracleLDDetector(nn.Module):
    """kd_one_stage.py:12-108: student (grad) + plain-attribute teacher (no grad) + LDHead loss."""
    strides = (8, 16, 32, 64, 128)

    def __init__(self, student_depth=50, teacher_depth=101, loss_cfg=None, imitation='finegrained', v2=False):
        """v2: LDv2Head student + GFocalHead teacher (configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py)."""
        super().__init__()
        self.v2 = v2
        s = OracleGFL(student_depth, v2)
        self.backbone, self.neck, self.bbox_head = s.backbone, s.neck, s.bbox_head
        object.__setattr__(self, 'teacher_model', OracleGFL(teacher_depth, v2))
        self.teacher_model.eval()
        self.loss_cfg = dict(DEFAULT_LOSS_CFG, **(loss_cfg or {}))
        self.imitation = imitation

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels):
        x = self.neck(self.backbone(img))
        with torch.no_grad():
            tx = self.teacher_model.extract_feat(img)
            t_out = self.teacher_model.bbox_head(tx)
        out = self.bbox_head(x)
        cls, reg = out[0], out[1]
        sizes = [tuple(c.shape[-2:]) for c in cls]
        tg = ld_targets(sizes, self.strides, img_metas, gt_bboxes, gt_labels,
                        imitation='finegrained' if self.imitation == 'gibox' else self.imitation)
        if self.v2:   # ld_gflv2.py:327: soft_teacher = (teacher cls_score [unused], bbox_pred, cls_feat)
            return ld_loss(cls, reg, t_out[2], t_out[1], x, tx, tg, self.strides, self.loss_cfg,
                           imitation=self.imitation, cls_feats=out[2], num_classes=80)
        return ld_loss(cls, reg, t_out[0], t_out[1], x, tx, tg, self.strides, self.loss_cfg,
                       imitation=self.imitation)

    def step(self, batch, backward=True):
        for p in self.parameters():
            p.grad = None
        losses = self.forward_train(batch['img'], batch['img_metas'], batch['gt_bboxes'],
                                    batch['gt_labels'])
        loss, log = parse_losses(losses)
        if backward:
            loss.backward()
        return losses, loss, log

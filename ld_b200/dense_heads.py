"""GFLHead / LDHead registered under the reference's names with the reference's constructor kwargs,
parameter names and call signatures, executing on the sm_100a kernels.

Reference: mmdet/models/dense_heads/anchor_head.py:36-96, gfl_head.py:15-194, ld_head.py:43-638.
The per-level Python loops of the reference (`multi_apply(forward_single)`, `multi_apply(loss_single)`,
`multi_apply(_get_target_single)`) collapse into: tower convs (tcgen05) + GroupNorm kernels per level,
ONE target-generation call for the whole batch, ONE pre-pass and ONE fused loss launch for all five
levels (row space).
"""
import dataclasses
import math
import os

import torch
from torch import nn

from . import engine, lib as L, ops
from .anchors import AnchorGenerator  # noqa: F401  (registers)
from .backbones import _Conv2d, nchw_view, nhwc, rows_of
from .registry import (HEADS, build_anchor_generator, build_assigner, build_bbox_coder, build_loss,
                       build_sampler)


class _ConvGN(nn.Module):
    """mmcv ConvModule(conv no-bias -> GN -> ReLU): parameters `.conv.weight`, `.gn.{weight,bias}`."""

    def __init__(self, cin, cout, norm_cfg):
        super().__init__()
        self.conv = _Conv2d(cin, cout, 3, 1, 1, bias=False)
        self.gn = nn.GroupNorm(norm_cfg.get('num_groups', 32), cout, eps=norm_cfg.get('eps', 1e-5))
        for p in self.gn.parameters():
            p.requires_grad = norm_cfg.get('requires_grad', True)

    def forward(self, x):
        c = engine.conv_affine_act(x, self.conv._cl(), None, None, None, 1, 1, False)
        return engine.gn_relu(c, self.gn.weight, self.gn.bias, self.gn.num_groups, self.gn.eps)


class Scale(nn.Module):
    """mmcv.cnn.Scale."""

    def __init__(self, scale=1.0):
        super().__init__()
        self.scale = nn.Parameter(torch.tensor(scale, dtype=torch.float))

    def forward(self, x):
        return x * self.scale


class Integral(nn.Module):
    """gfl_head.py:15-44 (buffer kept for state_dict parity; the expectation is computed in-kernel)."""

    def __init__(self, reg_max=16):
        super().__init__()
        self.reg_max = reg_max
        self.register_buffer('project', torch.linspace(0, reg_max, reg_max + 1))


@HEADS.register_module()
class GFLHead(nn.Module):
    def __init__(self, num_classes, in_channels, stacked_convs=4, conv_cfg=None,
                 norm_cfg=dict(type='GN', num_groups=32, requires_grad=True),
                 loss_dfl=dict(type='DistributionFocalLoss', loss_weight=0.25), reg_max=16,
                 feat_channels=256,
                 anchor_generator=dict(type='AnchorGenerator', scales=[8, 16, 32], ratios=[0.5, 1.0, 2.0],
                                       strides=[4, 8, 16, 32, 64]),
                 bbox_coder=dict(type='DeltaXYWHBBoxCoder', clip_border=True, target_means=(.0, .0, .0, .0),
                                 target_stds=(1.0, 1.0, 1.0, 1.0)),
                 reg_decoded_bbox=False,
                 loss_cls=dict(type='CrossEntropyLoss', use_sigmoid=True, loss_weight=1.0),
                 loss_bbox=dict(type='SmoothL1Loss', beta=1.0 / 9.0, loss_weight=1.0),
                 train_cfg=None, test_cfg=None):
        super().__init__()
        # ---- AnchorHead.__init__ (anchor_head.py:36-96)
        self.in_channels, self.num_classes, self.feat_channels = in_channels, num_classes, feat_channels
        self.use_sigmoid_cls = loss_cls.get('use_sigmoid', False)
        self.cls_out_channels = num_classes if self.use_sigmoid_cls else num_classes + 1
        if self.cls_out_channels <= 0:
            raise ValueError(f'num_classes={num_classes} is too small')
        self.reg_decoded_bbox = reg_decoded_bbox
        self.bbox_coder = build_bbox_coder(bbox_coder)
        self.loss_cls = build_loss(loss_cls)
        self.loss_bbox = build_loss(loss_bbox)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        if self.train_cfg:
            self.assigner = build_assigner(self.train_cfg.assigner)
            self.sampler = build_sampler(dict(type='PseudoSampler'), context=self)
        self.anchor_generator = build_anchor_generator(anchor_generator)
        self.num_anchors = self.anchor_generator.num_base_anchors[0]
        # ---- GFLHead.__init__ (gfl_head.py:77-100)
        self.stacked_convs, self.conv_cfg, self.norm_cfg, self.reg_max = stacked_convs, conv_cfg, norm_cfg, reg_max
        if conv_cfg is not None or norm_cfg.get('type') != 'GN':
            raise NotImplementedError('ld_b200 GFLHead implements conv_cfg=None + GN towers (all GFL/LD configs)')
        self.integral = Integral(reg_max)
        self.loss_dfl = build_loss(loss_dfl)
        self._init_layers()

    def _init_layers(self):
        self.cls_convs = nn.ModuleList()
        self.reg_convs = nn.ModuleList()
        for i in range(self.stacked_convs):
            chn = self.in_channels if i == 0 else self.feat_channels
            self.cls_convs.append(_ConvGN(chn, self.feat_channels, self.norm_cfg))
            self.reg_convs.append(_ConvGN(chn, self.feat_channels, self.norm_cfg))
        assert self.num_anchors == 1, 'anchor free version'
        self.gfl_cls = _Conv2d(self.feat_channels, self.cls_out_channels, 3, padding=1)
        self.gfl_reg = _Conv2d(self.feat_channels, 4 * (self.reg_max + 1), 3, padding=1)
        self.scales = nn.ModuleList([Scale(1.0) for _ in self.anchor_generator.strides])

    def init_weights(self):
        """gfl_head.py:135-143."""
        for m in list(self.cls_convs) + list(self.reg_convs):
            nn.init.normal_(m.conv.weight, 0, 0.01)
        nn.init.normal_(self.gfl_cls.weight, 0, 0.01)
        nn.init.constant_(self.gfl_cls.bias, float(-math.log((1 - 0.01) / 0.01)))
        nn.init.normal_(self.gfl_reg.weight, 0, 0.01)
        nn.init.constant_(self.gfl_reg.bias, 0)

    # ------------------------------------------------------------------ forward
    def towers(self, feats):
        """feats: NCHW-shaped level tensors -> per-level (cls_feat, reg_feat) NHWC bf16."""
        xs = [nhwc(x) for x in feats]
        if len(xs) > 5:
            raise NotImplementedError('more than 5 pyramid levels')
        # the two towers are independent chains: the regression tower runs on a second stream so that its
        # (bandwidth-bound, small-smem) GroupNorm kernels overlap the other tower's tensor-core convs; autograd
        # replays each backward op on its forward stream, so the backward overlaps the same way
        cfs, rfs = xs, xs
        use_side = xs[0].is_cuda and os.environ.get('LDB200_NO_TOWER_STREAM') is None
        main = torch.cuda.current_stream() if use_side else None
        side = self._tower_stream() if use_side else None
        # one tower layer = one node: conv of all levels (1 launch, GroupNorm statistics from its epilogue) +
        # normalise/ReLU of all levels (1 launch).  LDB200_GN_PER_LEVEL=1: the per-level GroupNorm launches (A/B).
        fused = os.environ.get('LDB200_GN_PER_LEVEL') is None and xs[0].shape[-1] % 32 == 0

        def layer(m, fs):
            if fused and m.gn.num_groups * 8 == m.conv.out_channels:
                return engine.conv_gn_relu_levels(m.conv._cl(), m.gn.weight, m.gn.bias, fs, 1, m.gn.num_groups, m.gn.eps)
            cs = engine.conv_levels(m.conv._cl(), fs)
            return [engine.gn_relu(c, m.gn.weight, m.gn.bias, m.gn.num_groups, m.gn.eps) for c in cs]
        if use_side:
            side.wait_stream(main)
            with torch.cuda.stream(side):
                for mr in self.reg_convs:
                    rfs = layer(mr, rfs)
        for mc in self.cls_convs:
            cfs = layer(mc, cfs)
        if use_side:
            main.wait_stream(side)
            for r in rfs:
                r.record_stream(main)
        else:
            for mr in self.reg_convs:
                rfs = layer(mr, rfs)
        return list(zip(cfs, rfs))

    def _tower_stream(self):
        dev = torch.cuda.current_device()
        st = self.__dict__.get('_side_stream')
        if st is None or st.device.index != dev:
            st = torch.cuda.Stream(device=dev)
            self.__dict__['_side_stream'] = st
        return st

    def final_convs(self, tower_out, cls_rows, reg_rows):
        """gfl_cls / gfl_reg*scale (gfl_head.py:181-182) written as fp32 NHWC rows (row space)."""
        o = 0
        wc, wr = engine.packed_weight(self.gfl_cls._cl()), engine.packed_weight(self.gfl_reg._cl())
        bc = self.gfl_cls.bias.detach()
        nb = 4 * (self.reg_max + 1)
        cls_outs = []
        for (cf, rf), sc in zip(tower_out, self.scales):
            n, h, w, _ = cf.shape
            cnt = n * h * w
            s = sc.scale.detach()
            cls_outs.append(cls_rows[o:o + cnt].view(n, h, w, -1))
            ops.conv2d_tc(rf, wr, nb, 3, 3, 1, 1, out=reg_rows[o:o + cnt].view(n, h, w, -1),
                          scale=s.expand(nb).contiguous(), bias=(self.gfl_reg.bias.detach() * s).contiguous())
            o += cnt
        # gfl_cls shares weights AND epilogue (bias) across levels: one launch for the whole pyramid
        ops.conv2d_tc_multi([t[0] for t in tower_out], wc, self.cls_out_channels, 3, 3, 1, 1, outs=cls_outs,
                            bias=bc.contiguous())

    def alloc_rows(self, feats):
        n = feats[0].shape[0]
        sizes = [tuple(f.shape[2:]) for f in feats]
        R = n * sum(h * w for h, w in sizes)
        dev = feats[0].device
        return (torch.empty(R, self.cls_out_channels, dtype=torch.float32, device=dev),
                torch.empty(R, 4 * (self.reg_max + 1), dtype=torch.float32, device=dev), sizes)

    def get_bboxes(self, cls_scores, bbox_preds, img_metas, cfg=None, rescale=False, with_nms=True):
        return _get_bboxes(self, cls_scores, bbox_preds, img_metas, cfg, rescale, with_nms)

    def forward(self, feats):
        """gfl_head.py:145-159: -> (cls_scores, bbox_preds), lists of NCHW-shaped fp32 level tensors.
        Inference-style forward (no autograd through the two output convs); training goes through
        `LDHead.forward_train`, which fuses them with the loss."""
        with torch.no_grad():
            cls_rows, reg_rows, sizes = self.alloc_rows(feats)
            self.final_convs(self.towers(feats), cls_rows, reg_rows)
        pyr = L.Pyramid.make(feats[0].shape[0], sizes, [s[0] for s in self.anchor_generator.strides])
        return ops.from_rows(cls_rows, pyr), ops.from_rows(reg_rows, pyr)


def _get_bboxes(head, cls_scores, bbox_preds, img_metas, cfg=None, rescale=False, with_nms=True):
    """`AnchorHead.get_bboxes` -> `GFLHead._get_bboxes` (anchor_head.py:518-600, gfl_head.py:354-451) for the whole batch
    in four launches (ldb200_gfl_get_bboxes): -> list of (det_bboxes [k, 5], det_labels [k]) per image."""
    cfg = head.test_cfg if cfg is None else cfg
    if not with_nms:
        raise NotImplementedError('with_nms=False (only used by aug-test merging, outside the LD path)')
    nms = cfg.get('nms', dict(type='nms', iou_threshold=0.6)) if hasattr(cfg, 'get') else cfg['nms']
    if nms.get('type', 'nms') != 'nms':
        raise NotImplementedError(f"nms type {nms.get('type')!r}: the kernel implements type='nms' (every GFL / LD test_cfg); "
                                  f"'voting_cluster_diounms' (bbox_nms.py:148-181) is not built")
    n = cls_scores[0].shape[0]
    sizes = [tuple(c.shape[2:]) for c in cls_scores]
    pyr = L.Pyramid.make(n, sizes, [s[0] for s in head.anchor_generator.strides])
    cls_rows = rows_of([c.detach().float() for c in cls_scores]).contiguous()
    reg_rows = rows_of([b.detach().float() for b in bbox_preds]).contiguous()
    img_hw = [m['img_shape'][:2] for m in img_metas]
    sf = None
    if rescale:
        import numpy as np
        sf = [np.broadcast_to(np.asarray(m['scale_factor'], dtype=np.float32), (4,)) for m in img_metas]
    dets, labels, counts = ops.gfl_get_bboxes(pyr, cls_rows, reg_rows, img_hw, sf, nms_pre=cfg.get('nms_pre', -1),
                                              score_thr=cfg['score_thr'], iou_thr=nms['iou_threshold'],
                                              max_per_img=cfg['max_per_img'], use_sigmoid=head.use_sigmoid_cls)
    counts = counts.cpu()
    if bool((counts < 0).any()):
        raise RuntimeError('ldb200_gfl_get_bboxes: more than 16384 (box, class) pairs above score_thr in one image')
    return [(dets[i, :int(counts[i])], labels[i, :int(counts[i])].long()) for i in range(n)]


# region mask produced by the target pass; 'gibox' rows come from `ops.gi_region` at loss time instead (the mask
# of the target pass is then unused)
IM_MODES = {'finegrained': L.IM_FINEGRAINED, 'fitnet': L.IM_CENTER, 'gibox': L.IM_FINEGRAINED}


def _upstream_term_grad(head, g_losses, nl):
    """The fused kernels emit the input gradients in the FORWARD pass for upstream weight 1 on every (term, level) --
    what `BaseDetector._parse_losses` (sum of all terms) produces.  Any other upstream gradient (loss scaling, a
    re-weighted or partial sum of the returned terms) is honoured by re-running the gradient pass with
    `term_grad = g_losses`, which the kernel supports per term and level.  Returns None when the precomputed gradients
    are valid, else the [8, MAX_LEVELS] host tensor to re-run with.  The check costs one device->host read, so it is
    skipped during CUDA-graph capture and when the caller vouches for the plain sum (`LDTrainStep`)."""
    if g_losses is None or getattr(head, '_unit_loss_grad', False) or \
            (g_losses.is_cuda and torch.cuda.is_current_stream_capturing()):
        return None
    g = g_losses[:L.NUM_TERMS - 1].detach().float().cpu()
    if bool((g[:, :nl] == 1).all()):
        return None
    return g.contiguous()


class _LDLossFn(torch.autograd.Function):
    """gfl_cls/gfl_reg output convs + LDHead.loss (ld_head.py:284-375) as one autograd node.

    inputs: per-level tower outputs (cls_feat, reg_feat), per-level student FPN maps, the four output
    conv parameters and the per-level Scale parameters.  forward computes the losses AND all input
    gradients (the fused kernel emits d_cls/d_reg/d_x in the same launch, upstream weights = 1 as in
    BaseDetector._parse_losses); backward hands them out."""

    @staticmethod
    def forward(ctx, head, aux, *tensors):
        nl = aux['num_levels']
        cls_feats, reg_feats, xs = tensors[:nl], tensors[nl:2 * nl], tensors[2 * nl:3 * nl]
        w_cls, b_cls, w_reg, b_reg = tensors[3 * nl:3 * nl + 4]
        scales = tensors[3 * nl + 4:]
        dev = cls_feats[0].device
        C, NB = head.cls_out_channels, 4 * (head.reg_max + 1)
        x_rows = rows_of([nchw_view(x.detach()) for x in xs])
        pyr = aux['pyr']
        R = pyr.rows
        cls_rows = torch.empty(R, C, dtype=torch.float32, device=dev)
        reg_rows = torch.empty(R, NB, dtype=torch.float32, device=dev)
        head.final_convs(list(zip(cls_feats, reg_feats)), cls_rows, reg_rows)
        tg = aux['targets']
        cfg = aux['loss_cfg']
        wsum, _ = ops.ld_loss_prepass(cfg, cls_rows, reg_rows, tg)
        norm = ops.loss_normalisers(tg, wsum)
        dcls_ld, dreg_ld = (C + 7) // 8 * 8, (NB + 7) // 8 * 8
        tg_im, max_fg = tg, aux.get('max_fg')
        if not aux['im_active']:
            max_fg = 1      # IM weight 0: no imitation rows are written (the kernel skips phase 3), keep d_im a stub
        if head.imitation_method == 'gibox' and aux['im_active']:
            # ld_head.py:170-176: imitate on the General-Instance rows (GI-box NMS picks, 10 per level) instead of a
            # target-derived region; they depend on this iteration's student and teacher head outputs
            gi_rows, gi_counts = ops.gi_region(cfg, cls_rows, reg_rows, aux['t_cls_rows'], aux['t_reg_rows'], tg,
                                               iou_thr=head.gi_iou_thr, topk=head.gi_topk)
            tg_im = dataclasses.replace(tg, fg_rows=gi_rows, counts=gi_counts)
            max_fg = gi_rows.numel()
        aux['targets_im'] = tg_im

        def run(term_grad):
            return ops.ld_loss_fused(
                cfg, cls_rows, reg_rows, aux['t_cls_rows'], aux['t_reg_rows'], x_rows, aux['t_x_rows'], tg_im, norm,
                term_grad=term_grad, grad_dtype=xs[0].dtype, max_fg=max_fg, d_cls_ld=dcls_ld, d_reg_ld=dreg_ld)
        losses, d_cls, d_reg, d_im = run(aux['term_grad'])
        ctx.regrad = run
        ctx.head, ctx.aux = head, aux
        ctx.save_for_backward(*tensors, cls_rows, reg_rows, d_cls, d_reg, d_im)
        ctx.mark_non_differentiable(norm)
        aux['cls_rows'], aux['reg_rows'] = cls_rows, reg_rows
        return losses, norm

    @staticmethod
    def backward(ctx, g_losses, _g_norm):
        head, aux = ctx.head, ctx.aux
        nl = aux['num_levels']
        saved = ctx.saved_tensors
        tensors, (cls_rows, reg_rows, d_cls, d_reg, d_im) = saved[:-5], saved[-5:]
        cls_feats, reg_feats, xs = tensors[:nl], tensors[nl:2 * nl], tensors[2 * nl:3 * nl]
        w_cls, b_cls, w_reg, b_reg = tensors[3 * nl:3 * nl + 4]
        scales = tensors[3 * nl + 4:]
        C, NB = head.cls_out_channels, 4 * (head.reg_max + 1)
        tg = aux['targets']
        tgrad = _upstream_term_grad(head, g_losses, nl)
        if tgrad is not None:     # non-unit upstream gradient: redo the gradient pass with it
            _, d_cls, d_reg, d_im = ctx.regrad(tgrad)
        dev = d_cls.device
        # conv biases: column sums of the row-space gradients
        g_bc = torch.zeros(d_cls.shape[1], dtype=torch.float32, device=dev)
        ops.colsum(d_cls, g_bc)
        g_bc = g_bc[:C]
        ci = w_cls.shape[1]
        Gc = torch.zeros(C, 9 * ci, dtype=torch.float32, device=dev)
        g_wr = torch.zeros(NB, 9 * ci, dtype=torch.float32, device=dev)
        g_br = torch.zeros(NB, dtype=torch.float32, device=dev)
        g_scales, g_cf, g_rf, g_x = [], [], [], []
        o = 0
        x_dense = None
        if aux['im_active']:
            x_dense = torch.zeros(tg.pyr.rows, xs[0].shape[-1], dtype=xs[0].dtype, device=dev)
            tg_im = aux['targets_im']
            ops.scatter_add_rows(x_dense, d_im, tg_im.fg_rows, tg_im.counts[1:2])
        for l in range(nl):
            n, h, w, _ = cls_feats[l].shape
            cnt = n * h * w
            gc = d_cls[o:o + cnt].view(n, h, w, -1)[..., :C]
            gr = d_reg[o:o + cnt].view(n, h, w, -1)[..., :NB]
            s = scales[l].detach()
            # bbox_pred = scale_l * (conv(reg_feat) + b)  (gfl_head.py:182):
            #   d scale_l = sum d_reg * bbox_pred / scale_l ; dW += scale_l * wgrad ; db += scale_l * colsum
            g_scales.append((gr.float() * reg_rows[o:o + cnt].view(n, h, w, NB)).sum() / s)
            ops.conv2d_wgrad_tc(gc, cls_feats[l], Gc, 3, 3, 1, 1)
            Gr = torch.zeros(NB, 9 * ci, dtype=torch.float32, device=dev)
            ops.conv2d_wgrad_tc(gr, reg_feats[l], Gr, 3, 3, 1, 1)
            g_wr += Gr * s
            blk = torch.zeros(d_reg.shape[1], dtype=torch.float32, device=dev)
            ops.colsum(d_reg[o:o + cnt], blk)
            g_br += blk[:NB] * s
            g_cf.append(gc)
            g_rf.append(engine.conv_dgrad(gr, w_reg, s.expand(NB), 1, 1, (h, w)))
            g_x.append(None if x_dense is None else x_dense[o:o + cnt].view(n, h, w, -1))
            o += cnt
        wd, cop = engine.dgrad_weight(w_cls, None)
        taps = [(1 - s_, 1 - r_, (r_ * 3 + s_) * cop) for r_ in range(3) for s_ in range(3)]
        g_cf = ops.conv2d_tc_multi(g_cf, wd, ci, 3, 3, 1, 1, taps=taps)
        g_wc = Gc.view(C, 3, 3, ci).permute(0, 3, 1, 2)
        g_wr = g_wr.view(NB, 3, 3, ci).permute(0, 3, 1, 2)
        return (None, None, *g_cf, *g_rf, *g_x, g_wc, g_bc, g_wr, g_br, *g_scales)


@HEADS.register_module()
class LDHead(GFLHead):
    """ld_head.py:43-71: GFLHead + localization distillation (LD), VLR LD, classification KD and
    feature imitation losses."""

    def __init__(self, num_classes, in_channels,
                 loss_ld=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10),
                 loss_ld_vlr=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10),
                 loss_kd=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=10, T=2),
                 loss_im=dict(type='IMLoss', loss_weight=0), imitation_method='gibox', **kwargs):
        super().__init__(num_classes, in_channels, **kwargs)
        self.imitation_method = imitation_method
        self.gi_iou_thr, self.gi_topk = 0.3, 10    # ld_head.py:637 (hard-coded in the reference)
        self.loss_im = build_loss(loss_im)
        self.loss_ld = build_loss(loss_ld)
        self.loss_ld_vlr = build_loss(loss_ld_vlr)
        self.loss_kd = build_loss(loss_kd)
        from .registry import build_iou_calculator
        self.iou_calculator = build_iou_calculator(dict(type='BboxOverlaps2D'))

    # ------------------------------------------------------------------ config -> kernel parameters
    def loss_cfg(self):
        if type(self.loss_bbox).__name__ != 'GIoULoss' or type(self.loss_cls).__name__ != 'QualityFocalLoss':
            raise NotImplementedError('the fused LD loss kernel implements QualityFocalLoss + GIoULoss '
                                      '(every configs/ld/*gflv1* file)')
        if self.loss_cls.use_sigmoid == self._qfl_on_scores:
            raise NotImplementedError(f'{type(self).__name__} needs QualityFocalLoss(use_sigmoid='
                                      f'{not self._qfl_on_scores}) (LDHead: logits, LDv2Head: joint scores)')
        return ops.make_loss_cfg(
            num_classes=self.cls_out_channels, reg_max=self.reg_max, feat_channels=self.in_channels,
            qfl_weight=self.loss_cls.loss_weight, qfl_beta=self.loss_cls.beta,
            giou_weight=self.loss_bbox.loss_weight, giou_eps=self.loss_bbox.eps,
            dfl_weight=self.loss_dfl.loss_weight, ld_weight=self.loss_ld.loss_weight, ld_T=self.loss_ld.T,
            ld_vlr_weight=self.loss_ld_vlr.loss_weight, ld_vlr_T=self.loss_ld_vlr.T,
            kd_weight=self.loss_kd.loss_weight, kd_T=self.loss_kd.T, im_weight=self.loss_im.loss_weight)

    _static_targets = None   # set by LDTrainStep.capture(): targets are produced outside the CUDA graph
    _qfl_on_scores = False   # LDv2Head: QFL runs on joint scores (use_sigmoid=False)

    def get_targets(self, featmap_sizes, img_metas, gt_bboxes, gt_labels, device, static=True):
        """ld_head.py:377-447 for the whole batch in one call (row space)."""
        if static and self._static_targets is not None and gt_bboxes is None:
            return self._static_targets
        if self.imitation_method not in IM_MODES:
            # 'decouple' (ld_head.py:177-185) pairs x[non-region rows] with teacher_x[region rows] in one mse_loss:
            # the reference itself fails on the shape mismatch unless the two row counts happen to coincide
            raise NotImplementedError(f"imitation_method='{self.imitation_method}' is not one of {sorted(IM_MODES)} "
                                      f"(ld_head.py:170-191)")
        if getattr(self.train_cfg, 'allowed_border', -1) >= 0:
            raise NotImplementedError('allowed_border >= 0 (LD configs use -1)')
        pyr = L.Pyramid.make(len(img_metas), featmap_sizes, [s[0] for s in self.anchor_generator.strides])
        pad_hw = [m['pad_shape'][:2] for m in img_metas]
        return ops.ld_targets(pyr, gt_bboxes, gt_labels, pad_hw, topk=self.assigner.topk,
                              im_mode=IM_MODES[self.imitation_method], num_classes=self.num_classes, device=device)

    # ------------------------------------------------------------------ training entry points
    def forward_train(self, x, out_teacher, teacher_x, img_metas, gt_bboxes, gt_labels=None,
                      gt_bboxes_ignore=None, proposal_cfg=None, **kwargs):
        """ld_head.py:73-114.  x / teacher_x: tuples of NCHW-shaped FPN levels, out_teacher = (cls_scores,
        bbox_preds) of the teacher head."""
        assert (gt_labels is not None or self._static_targets is not None) and proposal_cfg is None
        if gt_bboxes_ignore is not None and any(g is not None and g.numel() for g in gt_bboxes_ignore):
            raise NotImplementedError('gt_bboxes_ignore (ignore_iof_thr=-1 in all LD configs)')
        sizes = [tuple(f.shape[2:]) for f in x]
        assert len(sizes) == self.anchor_generator.num_levels
        dev = x[0].device
        tg = self.get_targets(sizes, img_metas, gt_bboxes, gt_labels, dev)
        t_cls, t_reg = out_teacher
        aux = dict(num_levels=len(sizes), pyr=tg.pyr, targets=tg, loss_cfg=self.loss_cfg(),
                   t_cls_rows=rows_of([t.detach() for t in t_cls]).float(),
                   t_reg_rows=rows_of([t.detach() for t in t_reg]).float(),
                   t_x_rows=rows_of([t.detach() for t in teacher_x]),
                   term_grad=True, im_active=self.loss_im.loss_weight != 0)
        tow = self.towers(x)
        xs = [nhwc(f) for f in x]
        if aux['t_x_rows'].dtype != xs[0].dtype:
            aux['t_x_rows'] = aux['t_x_rows'].to(xs[0].dtype)
        args = [t[0] for t in tow] + [t[1] for t in tow] + xs + \
               [self.gfl_cls._cl(), self.gfl_cls.bias, self.gfl_reg._cl(), self.gfl_reg.bias] + \
               [s.scale for s in self.scales]
        losses, norm = _LDLossFn.apply(self, aux, *args)
        self._last_aux = aux
        nl = len(sizes)
        return {k: [losses[i, l] for l in range(nl)] for i, k in enumerate(ops.TERMS)}

    def loss(self, cls_scores, bbox_preds, gt_bboxes, gt_labels, soft_teacher, x, teacher_x, img_metas,
             gt_bboxes_ignore=None):
        """ld_head.py:284-375 on given head outputs (API parity; differentiable w.r.t. cls_scores,
        bbox_preds and x through the same fused kernel, gradients returned in fp32)."""
        sizes = [tuple(f.shape[2:]) for f in cls_scores]
        dev = cls_scores[0].device
        tg = self.get_targets(sizes, img_metas, gt_bboxes, gt_labels, dev)
        soft_label, soft_target = soft_teacher
        nl = len(sizes)
        losses = _LossOnlyFn.apply(self, tg, nl, *cls_scores, *bbox_preds, *x, *soft_label, *soft_target, *teacher_x)
        return {k: [losses[i, l] for l in range(nl)] for i, k in enumerate(ops.TERMS)}


class _LossOnlyFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, head, tg, nl, *t):
        cls, reg, xs, t_cls, t_reg, t_x = (t[i * nl:(i + 1) * nl] for i in range(6))
        cfg = head.loss_cfg()
        rows = [ops.to_rows([v.detach().float() for v in grp]) for grp in (cls, reg, t_cls, t_reg)]
        xr = ops.to_rows([v.detach() for v in xs])
        txr = ops.to_rows([v.detach() for v in t_x]).to(xr.dtype)
        if xr.dtype not in (torch.float32, torch.bfloat16):
            xr, txr = xr.float(), txr.float()
        wsum, _ = ops.ld_loss_prepass(cfg, rows[0], rows[1], tg)
        norm = ops.loss_normalisers(tg, wsum)
        tg_im, max_fg = tg, None
        if head.imitation_method == 'gibox' and head.loss_im.loss_weight != 0:   # ld_head.py:170-176
            gi_rows, gi_counts = ops.gi_region(cfg, rows[0], rows[1], rows[2], rows[3], tg, iou_thr=head.gi_iou_thr,
                                               topk=head.gi_topk)
            tg_im, max_fg = dataclasses.replace(tg, fg_rows=gi_rows, counts=gi_counts), gi_rows.numel()
        head._last_im_targets = tg_im    # (rows, counts) the imitation term used: inspected by the parity tests
        def run(term_grad):
            losses, d_cls, d_reg, d_im = ops.ld_loss_fused(cfg, rows[0], rows[1], rows[2], rows[3], xr, txr, tg_im, norm,
                                                           term_grad=term_grad, grad_dtype=torch.float32, max_fg=max_fg)
            d_x = torch.zeros(tg.pyr.rows, xr.shape[1], dtype=torch.float32, device=xr.device)
            ops.scatter_add_rows(d_x, d_im, tg_im.fg_rows, tg_im.counts[1:2])
            return losses, d_cls, d_reg, d_x
        losses, d_cls, d_reg, d_x = run(True)
        ctx.pyr, ctx.nl, ctx.head, ctx.regrad = tg.pyr, nl, head, run
        ctx.save_for_backward(d_cls, d_reg, d_x)
        return losses

    @staticmethod
    def backward(ctx, g):
        d_cls, d_reg, d_x = ctx.saved_tensors
        nl = ctx.nl
        tgrad = _upstream_term_grad(ctx.head, g, nl)
        if tgrad is not None:
            _, d_cls, d_reg, d_x = ctx.regrad(tgrad)
        out = [None, None, None]
        for d in (d_cls, d_reg, d_x):
            out += [v for v in ops.from_rows(d, ctx.pyr)]
        out += [None] * (3 * nl)
        return tuple(out)


# ---------------------------------------------------------------------------------------------------------
# GFLv2 / LDv2: registry surface, parameters and checkpoint keys of the reference classes on the
# kernels `ldb200_dgqp_forward/_backward` (quality branch, dgqp.cu), `ldb200_gi_region_v2` (raw-score GI boxes) and the
# V2 instantiations of the prepass / fused loss (QFL on joint scores, KD on cls_feat).  GPU parity against the unmodified
# reference: tests/test_gpu_ldv2.py (goldens tests/golden/l2_*.pt, s2_*.pt; first GPU run in round 2).
@HEADS.register_module()
class GFocalHead(GFLHead):
    """gfocal_head.py:51-152: GFLHead + Distribution-Guided Quality Predictor (`reg_conf`)."""

    def __init__(self, num_classes, in_channels, reg_topk=4, reg_channels=64, add_mean=True, **kwargs):
        self.reg_topk, self.reg_channels, self.add_mean = reg_topk, reg_channels, add_mean
        self.total_dim = reg_topk + (1 if add_mean else 0)
        super().__init__(num_classes, in_channels, **kwargs)

    def _init_layers(self):
        super()._init_layers()
        # gfocal_head.py:147-152 (an nn.Sequential: checkpoint keys reg_conf.0.* and reg_conf.2.*)
        self.reg_conf = nn.Sequential(nn.Conv2d(4 * self.total_dim, self.reg_channels, 1), nn.ReLU(inplace=True),
                                      nn.Conv2d(self.reg_channels, 1, 1), nn.Sigmoid())

    def init_weights(self):
        super().init_weights()
        for m in self.reg_conf:      # gfocal_head.py:160-162: normal_init(std=0.01), bias 0
            if isinstance(m, nn.Conv2d):
                nn.init.normal_(m.weight, 0, 0.01)
                nn.init.constant_(m.bias, 0)

    def quality_params(self):
        """(W1 [hid, 4*total_dim], b1, W2 [hid], b2 [1]) in the layout `ops.dgqp_forward` takes."""
        c0, c2 = self.reg_conf[0], self.reg_conf[2]
        return (c0.weight.reshape(self.reg_channels, -1), c0.bias, c2.weight.reshape(-1), c2.bias)

    def forward(self, feats, with_scores=False):
        """gfocal_head.py:164-217 -> (cls_scores, bbox_preds, cls_feats) as lists of NCHW-shaped fp32 level tensors
        (inference-style, no autograd).  The joint scores cost one more launch and the LDv2 teacher's are never read
        (ld_gflv2.py:327), so they are None unless `with_scores`."""
        with torch.no_grad():
            feat_rows, reg_rows, sizes = self.alloc_rows(feats)
            self.final_convs(self.towers(feats), feat_rows, reg_rows)
            pyr = L.Pyramid.make(feats[0].shape[0], sizes, [s[0] for s in self.anchor_generator.strides])
            scores = None
            if with_scores:
                w1, b1, w2, b2 = [t.detach().float().contiguous() for t in self.quality_params()]
                _, score_rows = ops.dgqp_forward(reg_rows, feat_rows, w1, b1, w2, b2, self.reg_topk, self.add_mean)
                scores = ops.from_rows(score_rows, pyr)
        return scores, ops.from_rows(reg_rows, pyr), ops.from_rows(feat_rows, pyr)

    def forward_train(self, *args, **kwargs):
        raise NotImplementedError('plain GFLv2 training (GFocalHead.loss, gfocal_head.py:262-352) is not part of the LD '
                                  'hot path; LDv2Head.forward_train is')


@HEADS.register_module()
class LDv2Head(LDHead, GFocalHead):
    """ld_gflv2.py:43-71: GFocalHead + the LD / VLR-LD / KD / IM losses.  MRO LDv2Head -> LDHead -> GFocalHead ->
    GFLHead: the constructor signature, target assignment and loss configuration are LDHead's, the layers
    GFocalHead's."""
    _qfl_on_scores = True

    forward = GFocalHead.forward

    def forward_train(self, x, out_teacher, teacher_x, img_metas, gt_bboxes, gt_labels=None, gt_bboxes_ignore=None,
                      proposal_cfg=None, **kwargs):
        """ld_gflv2.py:73-114: out_teacher = GFocalHead.forward(teacher_x) = (scores [unused], bbox_preds, cls_feats)."""
        assert (gt_labels is not None or self._static_targets is not None) and proposal_cfg is None
        if gt_bboxes_ignore is not None and any(g is not None and g.numel() for g in gt_bboxes_ignore):
            raise NotImplementedError('gt_bboxes_ignore (ignore_iof_thr=-1 in all LD configs)')
        sizes = [tuple(f.shape[2:]) for f in x]
        assert len(sizes) == self.anchor_generator.num_levels
        tg = self.get_targets(sizes, img_metas, gt_bboxes, gt_labels, x[0].device)
        _, t_reg, t_feat = out_teacher
        aux = dict(num_levels=len(sizes), pyr=tg.pyr, targets=tg, loss_cfg=self.loss_cfg(),
                   t_cls_rows=rows_of([t.detach() for t in t_feat]).float(),
                   t_reg_rows=rows_of([t.detach() for t in t_reg]).float(),
                   t_x_rows=rows_of([t.detach() for t in teacher_x]),
                   term_grad=True, im_active=self.loss_im.loss_weight != 0)
        tow = self.towers(x)
        xs = [nhwc(f) for f in x]
        if aux['t_x_rows'].dtype != xs[0].dtype:
            aux['t_x_rows'] = aux['t_x_rows'].to(xs[0].dtype)
        c0, c2 = self.reg_conf[0], self.reg_conf[2]
        args = [t[0] for t in tow] + [t[1] for t in tow] + xs + \
               [self.gfl_cls._cl(), self.gfl_cls.bias, self.gfl_reg._cl(), self.gfl_reg.bias] + \
               [s.scale for s in self.scales] + [c0.weight, c0.bias, c2.weight, c2.bias]
        losses, norm = _LDv2LossFn.apply(self, aux, *args)
        self._last_aux = aux
        nl = len(sizes)
        return {k: [losses[i, l] for l in range(nl)] for i, k in enumerate(ops.TERMS)}

    def loss(self, cls_scores, bbox_preds, cls_feat, gt_bboxes, gt_labels, soft_teacher, x, teacher_x, img_metas,
             gt_bboxes_ignore=None):
        """ld_gflv2.py:286-375 on given head outputs: `cls_scores` = joint scores sigmoid(cls_feat) * quality,
        `soft_teacher` = (teacher joint scores [unused, ld_gflv2.py:327], teacher bbox_preds, teacher cls_feat)."""
        sizes = [tuple(f.shape[2:]) for f in cls_scores]
        dev = cls_scores[0].device
        tg = self.get_targets(sizes, img_metas, gt_bboxes, gt_labels, dev)
        _, soft_target, soft_label = soft_teacher
        nl = len(sizes)
        losses = _LossOnlyV2Fn.apply(self, tg, nl, *cls_scores, *bbox_preds, *cls_feat, *x, *soft_label, *soft_target,
                                     *teacher_x)
        return {k: [losses[i, l] for l in range(nl)] for i, k in enumerate(ops.TERMS)}


class _LossOnlyV2Fn(torch.autograd.Function):
    """`LDv2Head.loss` through ldb200_ld_loss_prepass_v2 / _gi_region_v2 / _ld_loss_fused_v2 (fp32 gradients)."""

    @staticmethod
    def forward(ctx, head, tg, nl, *t):
        score, reg, feat, xs, t_feat, t_reg, t_x = (t[i * nl:(i + 1) * nl] for i in range(7))
        cfg = head.loss_cfg()
        nfg = head.num_classes
        rows = [ops.to_rows([v.detach().float() for v in grp]) for grp in (score, reg, feat, t_feat, t_reg)]
        xr = ops.to_rows([v.detach() for v in xs])
        txr = ops.to_rows([v.detach() for v in t_x]).to(xr.dtype)
        if xr.dtype not in (torch.float32, torch.bfloat16):
            xr, txr = xr.float(), txr.float()
        wsum, _ = ops.ld_loss_prepass_v2(cfg, rows[0], rows[1], tg, nfg)
        norm = ops.loss_normalisers(tg, wsum)
        tg_im, max_fg = tg, None
        if head.imitation_method == 'gibox' and head.loss_im.loss_weight != 0:
            gi_rows, gi_counts = ops.gi_region_v2(cfg, rows[0], rows[1], rows[3], rows[4], tg, iou_thr=head.gi_iou_thr,
                                                  topk=head.gi_topk)
            tg_im, max_fg = dataclasses.replace(tg, fg_rows=gi_rows, counts=gi_counts), gi_rows.numel()
        head._last_im_targets = tg_im
        C = cfg.num_classes

        def run(term_grad):
            losses, d_score, d_kd, d_reg, d_im = ops.ld_loss_fused_v2(
                cfg, rows[0], rows[2], rows[1], rows[3], rows[4], xr, txr, tg_im, norm, nfg, term_grad=term_grad,
                grad_dtype=torch.float32, max_fg=max_fg)
            d_x = torch.zeros(tg.pyr.rows, xr.shape[1], dtype=torch.float32, device=xr.device)
            ops.scatter_add_rows(d_x, d_im, tg_im.fg_rows, tg_im.counts[1:2])
            return losses, (d_score[:, :C].contiguous(), d_reg, d_kd[:, :C].contiguous(), d_x)
        losses, grads = run(True)
        ctx.pyr, ctx.nl, ctx.head, ctx.regrad = tg.pyr, nl, head, run
        ctx.save_for_backward(*grads)
        return losses

    @staticmethod
    def backward(ctx, g):
        out = [None, None, None]
        grads = ctx.saved_tensors
        tgrad = _upstream_term_grad(ctx.head, g, ctx.nl)
        if tgrad is not None:
            grads = ctx.regrad(tgrad)[1]
        for d in grads:      # order of the inputs: scores, bbox_preds, cls_feat, x
            out += [v for v in ops.from_rows(d, ctx.pyr)]
        out += [None] * (3 * ctx.nl)
        return tuple(out)


class _LDv2LossFn(torch.autograd.Function):
    """`_LDLossFn` for LDv2Head: gfl_cls / gfl_reg output convs + DGQP quality branch (ldb200_dgqp_forward) +
    LDv2Head.loss (V2 prepass / GI region / fused loss) forward, and the whole gradient in the same pass: the fused
    kernel emits d_score / d_kd / d_reg / d_x, ldb200_dgqp_backward folds the quality branch back into d_cls_feat and
    d_reg and produces the reg_conf parameter gradients; `backward` hands them to the output convs' wgrad / dgrad.

    inputs: as `_LDLossFn` plus the four reg_conf tensors (conv1x1 weight/bias twice)."""

    @staticmethod
    def forward(ctx, head, aux, *tensors):
        nl = aux['num_levels']
        cls_feats, reg_feats, xs = tensors[:nl], tensors[nl:2 * nl], tensors[2 * nl:3 * nl]
        scales = tensors[3 * nl + 4:4 * nl + 4]
        q_w1, q_b1, q_w2, q_b2 = tensors[4 * nl + 4:]
        dev = cls_feats[0].device
        C, NB, nfg = head.cls_out_channels, 4 * (head.reg_max + 1), head.num_classes
        x_rows = rows_of([nchw_view(x.detach()) for x in xs])
        pyr = aux['pyr']
        R = pyr.rows
        feat_rows = torch.empty(R, C, dtype=torch.float32, device=dev)     # cls_feat logits
        reg_rows = torch.empty(R, NB, dtype=torch.float32, device=dev)
        head.final_convs(list(zip(cls_feats, reg_feats)), feat_rows, reg_rows)
        qp = [q_w1.detach().float().reshape(head.reg_channels, -1).contiguous(), q_b1.detach().float().contiguous(),
              q_w2.detach().float().reshape(-1).contiguous(), q_b2.detach().float().contiguous()]
        quality, score_rows = ops.dgqp_forward(reg_rows, feat_rows, *qp, reg_topk=head.reg_topk, add_mean=head.add_mean)
        tg = aux['targets']
        cfg = aux['loss_cfg']
        wsum, _ = ops.ld_loss_prepass_v2(cfg, score_rows, reg_rows, tg, nfg)
        norm = ops.loss_normalisers(tg, wsum)
        dcls_ld, dreg_ld = (C + 7) // 8 * 8, (NB + 7) // 8 * 8
        tg_im, max_fg = tg, aux.get('max_fg')
        if not aux['im_active']:
            max_fg = 1
        if head.imitation_method == 'gibox' and aux['im_active']:
            gi_rows, gi_counts = ops.gi_region_v2(cfg, score_rows, reg_rows, aux['t_cls_rows'], aux['t_reg_rows'], tg,
                                                  iou_thr=head.gi_iou_thr, topk=head.gi_topk)
            tg_im = dataclasses.replace(tg, fg_rows=gi_rows, counts=gi_counts)
            max_fg = gi_rows.numel()
        aux['targets_im'] = tg_im

        def run(term_grad):
            losses, d_score, d_kd, d_reg, d_im = ops.ld_loss_fused_v2(
                cfg, score_rows, feat_rows, reg_rows, aux['t_cls_rows'], aux['t_reg_rows'], x_rows, aux['t_x_rows'], tg_im,
                norm, nfg, term_grad=term_grad, grad_dtype=xs[0].dtype, max_fg=max_fg, d_cls_ld=dcls_ld,
                d_reg_ld=dreg_ld)
            d_feat, g_w1, g_b1, g_w2, g_b2 = ops.dgqp_backward(reg_rows, feat_rows, *qp, quality, d_score, d_reg,
                                                               d_kd=d_kd, reg_topk=head.reg_topk, add_mean=head.add_mean)
            return losses, (d_feat, d_reg, d_im, g_w1, g_b1, g_w2, g_b2)
        losses, (d_feat, d_reg, d_im, g_w1, g_b1, g_w2, g_b2) = run(aux['term_grad'])
        ctx.regrad = run
        ctx.head, ctx.aux = head, aux
        ctx.q_shapes = (tuple(q_w1.shape), tuple(q_w2.shape))
        ctx.save_for_backward(*tensors, feat_rows, reg_rows, d_feat, d_reg, d_im, g_w1, g_b1, g_w2, g_b2)
        ctx.mark_non_differentiable(norm)
        aux['cls_rows'], aux['reg_rows'], aux['score_rows'] = feat_rows, reg_rows, score_rows
        return losses, norm

    @staticmethod
    def backward(ctx, g_losses, _g_norm):
        head, aux = ctx.head, ctx.aux
        nl = aux['num_levels']
        saved = ctx.saved_tensors
        tensors, (feat_rows, reg_rows, d_cls, d_reg, d_im, g_w1, g_b1, g_w2, g_b2) = saved[:-9], saved[-9:]
        cls_feats, reg_feats, xs = tensors[:nl], tensors[nl:2 * nl], tensors[2 * nl:3 * nl]
        w_cls, b_cls, w_reg, b_reg = tensors[3 * nl:3 * nl + 4]
        scales = tensors[3 * nl + 4:4 * nl + 4]
        C, NB = head.cls_out_channels, 4 * (head.reg_max + 1)
        tg = aux['targets']
        tgrad = _upstream_term_grad(head, g_losses, nl)
        if tgrad is not None:
            d_cls, d_reg, d_im, g_w1, g_b1, g_w2, g_b2 = ctx.regrad(tgrad)[1]
        dev = d_cls.device
        g_bc = torch.zeros(d_cls.shape[1], dtype=torch.float32, device=dev)
        ops.colsum(d_cls, g_bc)
        g_bc = g_bc[:C]
        ci = w_cls.shape[1]
        Gc = torch.zeros(C, 9 * ci, dtype=torch.float32, device=dev)
        g_wr = torch.zeros(NB, 9 * ci, dtype=torch.float32, device=dev)
        g_br = torch.zeros(NB, dtype=torch.float32, device=dev)
        g_scales, g_cf, g_rf, g_x = [], [], [], []
        o = 0
        x_dense = None
        if aux['im_active']:
            x_dense = torch.zeros(tg.pyr.rows, xs[0].shape[-1], dtype=xs[0].dtype, device=dev)
            tg_im = aux['targets_im']
            ops.scatter_add_rows(x_dense, d_im, tg_im.fg_rows, tg_im.counts[1:2])
        for l in range(nl):
            n, h, w, _ = cls_feats[l].shape
            cnt = n * h * w
            gc = d_cls[o:o + cnt].view(n, h, w, -1)[..., :C]
            gr = d_reg[o:o + cnt].view(n, h, w, -1)[..., :NB]
            s = scales[l].detach()
            g_scales.append((gr.float() * reg_rows[o:o + cnt].view(n, h, w, NB)).sum() / s)   # gfl_head.py:182
            ops.conv2d_wgrad_tc(gc, cls_feats[l], Gc, 3, 3, 1, 1)
            Gr = torch.zeros(NB, 9 * ci, dtype=torch.float32, device=dev)
            ops.conv2d_wgrad_tc(gr, reg_feats[l], Gr, 3, 3, 1, 1)
            g_wr += Gr * s
            blk = torch.zeros(d_reg.shape[1], dtype=torch.float32, device=dev)
            ops.colsum(d_reg[o:o + cnt], blk)
            g_br += blk[:NB] * s
            g_cf.append(gc)
            g_rf.append(engine.conv_dgrad(gr, w_reg, s.expand(NB), 1, 1, (h, w)))
            g_x.append(None if x_dense is None else x_dense[o:o + cnt].view(n, h, w, -1))
            o += cnt
        wd, cop = engine.dgrad_weight(w_cls, None)
        taps = [(1 - s_, 1 - r_, (r_ * 3 + s_) * cop) for r_ in range(3) for s_ in range(3)]
        g_cf = ops.conv2d_tc_multi(g_cf, wd, ci, 3, 3, 1, 1, taps=taps)
        g_wc = Gc.view(C, 3, 3, ci).permute(0, 3, 1, 2)
        g_wr = g_wr.view(NB, 3, 3, ci).permute(0, 3, 1, 2)
        s1, s2 = ctx.q_shapes
        return (None, None, *g_cf, *g_rf, *g_x, g_wc, g_bc, g_wr, g_br, *g_scales,
                g_w1.reshape(s1), g_b1, g_w2.reshape(s2), g_b2)

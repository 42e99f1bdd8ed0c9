"""Generate `tests/golden/*.pt` by running the UNMODIFIED reference (through the mmcv stand-in)
on the seeded cases of `oracle/cases.py`.  Build container only (`/root/reference` must exist):

    python -m oracle.make_golden

The reference's own tests hold no golden vectors for this path, so these are
outputs of the reference code itself.  Fixtures are kept small (int16/int32 indices, fp32 values).
TEST INFRASTRUCTURE ONLY.
"""
import os
import sys

import numpy as np
import torch

from ld_b200 import synth
from . import cases, ref_harness as rh

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests', 'golden')


def ref_targets(head, n, h, w, gt_bboxes, gt_labels):
    """Reference `get_anchors` + `get_targets` (ld_head.py:319-336)."""
    sizes = cases.featmap_sizes(h, w)
    metas = synth.make_img_metas(n, h, w)
    anchor_list, valid_flag_list = head.get_anchors(sizes, metas, device='cpu')
    out = head.get_targets(anchor_list, valid_flag_list, gt_bboxes, metas,
                           gt_bboxes_ignore_list=None, gt_labels_list=gt_labels,
                           label_channels=head.cls_out_channels)
    (anchors, labels, label_w, bbox_t, _bbox_w, num_pos, _num_neg, vlr, im) = out
    return dict(anchors=anchors, labels=labels, label_weights=label_w, bbox_targets=bbox_t,
                vlr=vlr, im=im, num_total_pos=num_pos)


def gen_targets(head):
    for name, n, h, w, seed in cases.TARGET_CASES:
        gb, gl = synth.make_gts(n, h, w, seed)
        for mode in ('finegrained', 'fitnet'):
            head.imitation_method = mode
            t = ref_targets(head, n, h, w, gb, gl)
            if mode == 'finegrained':
                gold = dict(
                    labels=torch.cat(t['labels'], 1).to(torch.int16),
                    label_weights=torch.cat(t['label_weights'], 1).to(torch.uint8),
                    bbox_targets=torch.cat(t['bbox_targets'], 1),
                    vlr=torch.cat(t['vlr'], 1), im_finegrained=torch.cat(t['im'], 1).to(torch.uint8),
                    num_total_pos=int(t['num_total_pos']),
                    level_anchor0=[a[0, :2].clone() for a in t['anchors']])
            else:
                gold['im_center'] = torch.cat(t['im'], 1).to(torch.uint8)
        head.imitation_method = 'finegrained'
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, 'pos', gold['num_total_pos'], 'vlr', int((gold['vlr'] > 0).sum()),
              'im', int(gold['im_finegrained'].sum()))


def gen_losses(head):
    for name, n, h, w, seed in cases.LOSS_CASES:
        d = cases.loss_case_inputs(n, h, w, seed)
        leaves = {}
        for k in ('cls', 'reg', 'x'):
            leaves[k] = [t.clone().requires_grad_(True) for t in d[k]]
        losses = head.loss(leaves['cls'], leaves['reg'], d['gt_bboxes'], d['gt_labels'],
                           (d['t_cls'], d['t_reg']), leaves['x'], d['t_x'], d['img_metas'])
        total = sum(sum(v) for v in losses.values())
        total.backward()
        for t in leaves['x']:
            if t.grad is None:   # level without fine-grained rows: loss_im does not touch x
                t.grad = torch.zeros_like(t)
        gold = dict(losses={k: torch.stack([t.detach() for t in v]) for k, v in losses.items()},
                    total=total.detach(),
                    grad_cls=[t.grad for t in leaves['cls']],
                    grad_reg=[t.grad for t in leaves['reg']],
                    # grad_x is sparse (only fine-grained rows): keep per-level sums + norms
                    grad_x_norm=torch.stack([t.grad.norm() for t in leaves['x']]),
                    grad_x_rowsum=[t.grad.sum(1) for t in leaves['x']])
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, {k: [round(float(x), 5) for x in v] for k, v in gold['losses'].items()})


def gen_losses_gibox(head):
    """`imitation_method='gibox'` (the LDHead constructor default, ld_head.py:55): the reference's
    `get_gi_region` calls `.cuda()` on an index tensor (ld_head.py:628) -- on this CPU-only container
    `Tensor.cuda` is patched to the identity while the UNMODIFIED reference code runs, and the method is wrapped to
    record the indices it returns."""
    import torchvision  # noqa: F401  (registers torch.ops.torchvision.nms, which the reference calls directly)
    orig_cuda, orig_gi = torch.Tensor.cuda, head.get_gi_region
    picked = []

    def spy(*a, **k):
        idx = orig_gi(*a, **k)
        picked.append(idx.clone())
        return idx

    torch.Tensor.cuda = lambda self, *a, **k: self
    head.get_gi_region = spy
    head.imitation_method = 'gibox'
    try:
        for name, n, h, w, seed in cases.LOSS_CASES:
            d = cases.loss_case_inputs(n, h, w, seed)
            del picked[:]
            leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('cls', 'reg', 'x')}
            losses = head.loss(leaves['cls'], leaves['reg'], d['gt_bboxes'], d['gt_labels'],
                               (d['t_cls'], d['t_reg']), leaves['x'], d['t_x'], d['img_metas'])
            total = sum(sum(v) for v in losses.values())
            total.backward()
            for t in leaves['x']:
                if t.grad is None:   # level without positives: loss_im is zeroed (ld_head.py:251)
                    t.grad = torch.zeros_like(t)
            gold = dict(losses={k: torch.stack([t.detach() for t in v]) for k, v in losses.items()},
                        total=total.detach(), gi_idx=[p.to(torch.int32) for p in picked],
                        grad_x_norm=torch.stack([t.grad.norm() for t in leaves['x']]),
                        grad_x_rowsum=[t.grad.sum(1) for t in leaves['x']],
                        grad_cls_norm=torch.stack([t.grad.norm() for t in leaves['cls']]),
                        grad_reg_norm=torch.stack([t.grad.norm() for t in leaves['reg']]))
            torch.save(gold, os.path.join(OUT, name + '_gibox.pt'))
            print(name + '_gibox', [round(float(x), 5) for x in gold['losses']['loss_im']],
                  [p.tolist()[:4] for p in picked])
    finally:
        torch.Tensor.cuda, head.get_gi_region = orig_cuda, orig_gi
        head.imitation_method = 'finegrained'


def _with_cpu_gibox(fn):
    """Run `fn` with `Tensor.cuda` patched to the identity (ld_head.py:628 / ld_gflv2.py call `.cuda()` on an index
    tensor) and torchvision imported (registers torch.ops.torchvision.nms)."""
    import torchvision  # noqa: F401
    orig = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self
    try:
        return fn()
    finally:
        torch.Tensor.cuda = orig


def gen_losses_v2():
    """`LDv2Head.loss` (ld_gflv2.py:286-375) of configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py: gibox imitation, QFL on joint
    scores (use_sigmoid=False), KD on cls_feat, 81 class channels."""
    head = rh.reference_head(config_rel='configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py')
    assert type(head).__name__ == 'LDv2Head' and head.cls_out_channels == 81 and head.imitation_method == 'gibox'
    picked, orig_gi = [], head.get_gi_region

    def spy(*a, **k):
        idx = orig_gi(*a, **k)
        picked.append(idx.clone())
        return idx
    head.get_gi_region = spy
    for name, n, h, w, seed in cases.LOSS_CASES_V2:
        d = cases.loss_case_inputs_v2(n, h, w, seed)
        del picked[:]
        leaves = {k: [t.clone().requires_grad_(True) for t in d[k]] for k in ('score', 'cls_feat', 'reg', 'x')}

        def run():
            return head.loss(leaves['score'], leaves['reg'], leaves['cls_feat'], d['gt_bboxes'], d['gt_labels'],
                             (None, d['t_reg'], d['t_cls']), leaves['x'], d['t_x'], d['img_metas'])
        losses = _with_cpu_gibox(run)
        total = sum(sum(v) for v in losses.values())
        total.backward()
        for grp in leaves.values():
            for t in grp:
                if t.grad is None:
                    t.grad = torch.zeros_like(t)
        gold = dict(losses={k: torch.stack([t.detach() for t in v]) for k, v in losses.items()},
                    total=total.detach(), gi_idx=[p.to(torch.int32) for p in picked],
                    grad_score=[t.grad for t in leaves['score']],
                    grad_cls_feat_norm=torch.stack([t.grad.norm() for t in leaves['cls_feat']]),
                    grad_cls_feat_rowsum=[t.grad.sum(1) for t in leaves['cls_feat']],
                    grad_reg=[t.grad for t in leaves['reg']],
                    grad_x_rowsum=[t.grad.sum(1) for t in leaves['x']])
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, {k: [round(float(x), 5) for x in v] for k, v in gold['losses'].items()})


def gen_steps_v2():
    for name, sd, td, n, h, w, seed in cases.STEP_CASES_V2:
        cfg = rh.load_model_cfg('configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py')
        cfg['backbone']['depth'] = sd
        cfg['neck']['in_channels'] = [64, 128, 256, 512] if sd < 50 else [256, 512, 1024, 2048]
        cfg['teacher_config']['model']['backbone']['depth'] = td
        cfg['teacher_config']['model']['neck']['in_channels'] = \
            [64, 128, 256, 512] if td < 50 else [256, 512, 1024, 2048]
        m = rh.build_reference_detector(cfg)
        synth.seeded_init(m, seed)
        synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
        b = synth.make_batch(n, h, w, seed=1000 + seed)
        losses, loss, log = _with_cpu_gibox(lambda: rh.reference_step(m, b))
        named = dict(m.named_parameters())
        gold = dict(log_vars=log,
                    losses={k: torch.stack([t.detach() for t in v]) for k, v in losses.items()},
                    grad_stats={k: torch.tensor([float(p.grad.double().norm()), float(p.grad.double().sum()),
                                                 float(p.grad.abs().max())])
                                for k, p in named.items() if p.grad is not None},
                    no_grad=[k for k, p in named.items() if p.grad is None],
                    grads={k: named[k].grad.clone() for k in
                           ('bbox_head.reg_conf.0.weight', 'bbox_head.reg_conf.2.weight', 'bbox_head.gfl_cls.bias',
                            'bbox_head.gfl_reg.bias', 'bbox_head.scales.0.scale')})
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, {k: round(v, 5) for k, v in log.items()})


def gen_steps(case_list=None):
    for name, sd, td, n, h, w, seed in (case_list or cases.STEP_CASES):
        cfg = rh.load_model_cfg()
        cfg['backbone']['depth'] = sd
        cfg['neck']['in_channels'] = [64, 128, 256, 512] if sd < 50 else [256, 512, 1024, 2048]
        cfg['teacher_config']['model']['backbone']['depth'] = td
        cfg['teacher_config']['model']['neck']['in_channels'] = \
            [64, 128, 256, 512] if td < 50 else [256, 512, 1024, 2048]
        m = rh.build_reference_detector(cfg)
        synth.seeded_init(m, seed)
        synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
        b = synth.make_batch(n, h, w, seed=1000 + seed)
        losses, loss, log = rh.reference_step(m, b)
        gold = dict(log_vars=log,
                    losses={k: torch.stack([t.detach() for t in v]) for k, v in losses.items()},
                    grad_stats={k: torch.tensor([float(p.grad.double().norm()),
                                                 float(p.grad.double().sum()),
                                                 float(p.grad.abs().max())])
                                for k, p in m.named_parameters() if p.grad is not None},
                    no_grad=[k for k, p in m.named_parameters() if p.grad is None])
        # a few full gradients for direct comparison
        keep = ['bbox_head.gfl_cls.weight', 'bbox_head.gfl_reg.bias', 'bbox_head.scales.0.scale',
                'bbox_head.cls_convs.0.gn.weight', 'neck.lateral_convs.0.conv.bias',
                'backbone.layer4.0.bn1.weight' if sd >= 50 else 'backbone.layer4.0.bn1.weight']
        named = dict(m.named_parameters())
        gold['grads'] = {k: named[k].grad.clone() for k in keep}
        if case_list is not None:   # later fixtures: a fingerprint of EVERY gradient tensor (cases.grad_projection)
            gold['grad_proj'] = {k: cases.grad_projection(k, p.grad) for k, p in named.items() if p.grad is not None}
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, {k: round(v, 5) for k, v in log.items()})


def gen_curve():
    name, sd, td, n, h, w, seed, iters, base_lr, warm = cases.CURVE_CASE
    cfg = rh.load_model_cfg()
    cfg['backbone']['depth'] = sd
    cfg['neck']['in_channels'] = [64, 128, 256, 512]
    cfg['teacher_config']['model']['backbone']['depth'] = td
    cfg['teacher_config']['model']['neck']['in_channels'] = [64, 128, 256, 512]
    m = rh.build_reference_detector(cfg)
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    opt = torch.optim.SGD([p for p in m.parameters() if p.requires_grad], lr=base_lr, momentum=0.9, weight_decay=1e-4)
    curve, terms = [], []
    for it in range(iters):
        lr = base_lr * (1 - (1 - it / warm) * (1 - 0.001)) if it < warm else base_lr   # mmcv LrUpdaterHook 'linear'
        for gp in opt.param_groups:
            gp['lr'] = lr
        b = synth.make_batch(n, h, w, seed=2000 + it)
        losses, loss, log = rh.reference_step(m, b)
        opt.step()
        curve.append(float(loss.detach()))
        terms.append([log[k] for k in sorted(log) if k != 'loss'])
    gold = dict(loss=torch.tensor(curve), terms=torch.tensor(terms), term_names=sorted(k for k in log if k != 'loss'),
                final_checksum={k: float(p.detach().double().sum()) for k, p in list(m.named_parameters())[-6:]})
    torch.save(gold, os.path.join(OUT, name + '.pt'))
    print(name, [round(c, 4) for c in curve[:3]], '...', [round(c, 4) for c in curve[-3:]])


def gen_loss_modules():
    """QualityFocalLoss / DistributionFocalLoss / GIoULoss of the unmodified reference called stand-alone
    (gfocal_loss.py:77-179, iou_loss.py:325-360): values and d/d pred for every call variant of cases.LOSS_MODULE_CALLS."""
    from oracle import mmcv_standin
    mmcv_standin.import_reference(rh.REF_ROOT)
    from mmdet.models import build_loss
    mods = dict(qfl=build_loss(dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0)),
                qflp=build_loss(dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0)),
                dfl=build_loss(dict(type='DistributionFocalLoss', loss_weight=0.25)),
                giou=build_loss(dict(type='GIoULoss', loss_weight=2.0)))
    for name, rows, seed in cases.LOSS_MODULE_CASES:
        d = cases.loss_module_inputs(rows, seed)
        gold = {}
        for tag, kw in cases.LOSS_MODULE_CALLS:
            kw = dict(kw)
            w = None if kw.pop('no_weight', False) else d['weight']
            for key, pred, target in (('qfl', d['qfl_pred'], (d['label'], d['score'])),
                                      ('qflp', d['qfl_prob'], (d['label'], d['score'])),
                                      ('dfl', d['dfl_pred'], d['dfl_target']),
                                      ('giou', d['box_pred'], d['box_target'])):
                p = pred.clone().requires_grad_(True)
                out = mods[key](p, target, weight=w, **kw)
                # reduction='none': contract with a fixed seeded vector so that one backward covers every row
                up = torch.linspace(0.5, 1.5, out.numel()).reshape(out.shape) if out.dim() else torch.tensor(1.0)
                (out * up).sum().backward()
                gold[f'{key}_{tag}'] = dict(out=out.detach().clone(), grad=p.grad.clone())
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, {k: float(v['out'].sum()) for k, v in list(gold.items())[:4]})


def gen_infer():
    """`GFLHead.get_bboxes` of the unmodified reference head (anchor_head.py:518-600 -> gfl_head.py:354-451 ->
    bbox_nms.py:70-193; mmcv's batched_nms comes from the stand-in) on cases.INFER_CASES."""
    import numpy as np
    head = rh.reference_head()
    head.eval()
    for name, n, h, w, seed, rescale in cases.INFER_CASES:
        d = cases.infer_case_inputs(n, h, w, seed)
        metas = [dict(m, scale_factor=np.asarray(m['scale_factor'], dtype=np.float32)) for m in d['img_metas']]
        with torch.no_grad():
            res = head.get_bboxes(d['cls'], d['reg'], metas, rescale=rescale)
        torch.save(dict(dets=[r[0].clone() for r in res], labels=[r[1].clone() for r in res]), os.path.join(OUT, name + '.pt'))
        print(name, [tuple(r[0].shape) for r in res], [float(r[0][:, 4].max()) for r in res])


def gen_pipeline():
    """The reference's own transform CLASSES (Resize / RandomFlip / Normalize / Pad, transforms.py) run through the
    stand-in's cv2-backed mmcv functions on cases.PIPELINE_CASES; a sha256 of every output image (bit-exactness
    fingerprint), a strided sample of it, the transformed boxes and the meta keys are kept."""
    import hashlib
    from oracle import mmcv_standin, pipeline_ref as pr
    mmcv_standin.import_reference(rh.REF_ROOT)
    from mmdet.datasets.pipelines import transforms as T
    for name, sizes, seed, flips, img_scale in cases.PIPELINE_CASES:
        imgs, boxes, _ = pr.synth_images(sizes, seed)
        gold = dict(sha=[], sample=[], boxes=[], img_shape=[], pad_shape=[], scale_factor=[])
        for img, b, flip in zip(imgs, boxes, flips):
            r = dict(img=img, gt_bboxes=b.copy(), bbox_fields=['gt_bboxes'], img_fields=['img'], img_shape=img.shape,
                     ori_shape=img.shape, flip=bool(flip), flip_direction='horizontal' if flip else None)
            for t in (T.Resize(img_scale=img_scale, keep_ratio=True), T.RandomFlip(flip_ratio=0.5),
                      T.Normalize(mean=[123.675, 116.28, 103.53], std=[58.395, 57.12, 57.375], to_rgb=True),
                      T.Pad(size_divisor=32)):
                r = t(r)
            chw = np.ascontiguousarray(r['img'].transpose(2, 0, 1))
            gold['sha'].append(hashlib.sha256(chw.tobytes()).hexdigest())
            gold['sample'].append(torch.from_numpy(chw[:, ::7, ::5].copy()))
            gold['boxes'].append(torch.from_numpy(r['gt_bboxes'].copy()))
            gold['img_shape'].append(tuple(r['img_shape']))
            gold['pad_shape'].append(tuple(r['pad_shape']))
            gold['scale_factor'].append(torch.from_numpy(r['scale_factor'].copy()))
        torch.save(gold, os.path.join(OUT, name + '.pt'))
        print(name, gold['img_shape'], [s[:8] for s in gold['sha']])


def main():
    if not rh.reference_available():
        sys.exit('needs /root/reference (build container only)')
    os.makedirs(OUT, exist_ok=True)
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    if '--pipeline' in sys.argv:
        gen_pipeline()
        return
    if '--infer' in sys.argv:
        gen_infer()
        return
    if '--loss-modules' in sys.argv:
        gen_loss_modules()
        return
    if '--config1' in sys.argv:      # round 2: BASELINE configs[0] only (the other fixtures stay byte-identical)
        gen_steps([cases.STEP_CASE_CONFIG1])
        return
    head = rh.reference_head()
    gen_targets(head)
    gen_losses(head)
    gen_losses_gibox(head)
    gen_steps()
    gen_losses_v2()
    gen_steps_v2()
    gen_curve()


if __name__ == '__main__':
    main()

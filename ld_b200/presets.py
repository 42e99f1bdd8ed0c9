"""Model sections equivalent to the reference's LD configs, for environments where the reference tree
is absent (the B200 box).  `tests/test_config_dropin.py` checks (build container) that loading the
reference's own files through `Config.fromfile` yields the same dictionaries.

    ld_model('r50')  == configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py : model   (teacher = gfl_r101_fpn_mstrain_2x)
    ld_model('r18')  == configs/ld/ld_r18_gflv1_r101_fpn_coco_1x.py : model
"""
import copy

_CH = {18: [64, 128, 256, 512], 34: [64, 128, 256, 512], 50: [256, 512, 1024, 2048], 101: [256, 512, 1024, 2048]}

_TRAIN = dict(assigner=dict(type='ATSSAssigner', topk=9), allowed_border=-1, pos_weight=-1, debug=False)
_TEST = dict(nms_pre=1000, min_bbox_size=0, score_thr=0.05, nms=dict(type='nms', iou_threshold=0.6), max_per_img=100)
_ANCHORS = dict(type='AnchorGenerator', ratios=[1.0], octave_base_scale=8, scales_per_octave=1,
                strides=[8, 16, 32, 64, 128])


def _backbone(depth):
    return dict(type='ResNet', depth=depth, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=1,
                norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, style='pytorch')


def _neck(depth):
    return dict(type='FPN', in_channels=_CH[depth], out_channels=256, start_level=1, add_extra_convs='on_output',
                num_outs=5)


def gfl_model(depth=101):
    """configs/gfl/gfl_r50_fpn_1x_coco.py model with the backbone of gfl_r101_fpn_mstrain_2x_coco.py."""
    return dict(type='GFL', pretrained=f'torchvision://resnet{depth}', backbone=_backbone(depth), neck=_neck(depth),
                bbox_head=dict(type='GFLHead', num_classes=80, in_channels=256, stacked_convs=4, feat_channels=256,
                               anchor_generator=copy.deepcopy(_ANCHORS),
                               loss_cls=dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0),
                               loss_dfl=dict(type='DistributionFocalLoss', loss_weight=0.25), reg_max=16,
                               loss_bbox=dict(type='CIoULoss', loss_weight=2.0)),
                train_cfg=copy.deepcopy(_TRAIN), test_cfg=copy.deepcopy(_TEST))


def ld_model(student='r50', teacher_depth=101, main_ld_only=False):
    depth = int(student.lstrip('r'))
    head = dict(type='LDHead', num_classes=80, in_channels=256, stacked_convs=4, feat_channels=256,
                anchor_generator=copy.deepcopy(_ANCHORS),
                loss_cls=dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0),
                loss_dfl=dict(type='DistributionFocalLoss', loss_weight=0.25),
                loss_ld=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10),
                reg_max=16, loss_bbox=dict(type='GIoULoss', loss_weight=2.0))
    if depth >= 50 or True:
        # ld_r50_gflv1_r101_fpn_coco_1x.py:39-47 (the r18 file leaves these to the LDHead ctor defaults, which
        # select imitation_method='gibox'; 'finegrained' is what the R50 config trains with)
        head.update(loss_ld_vlr=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=0.25, T=10),
                    loss_kd=dict(type='KnowledgeDistillationKLDivLoss', loss_weight=10, T=2),
                    loss_im=dict(type='IMLoss', loss_weight=2.0), imitation_method='finegrained')
    if main_ld_only:   # BASELINE.json configs[2]: main-LD only
        head['loss_ld_vlr']['loss_weight'] = 0
        head['loss_kd']['loss_weight'] = 0
        head['loss_im']['loss_weight'] = 0
    return dict(type='KnowledgeDistillationSingleStageDetector', pretrained=None,
                teacher_config=dict(model=gfl_model(teacher_depth)), teacher_ckpt=None, output_feature=True,
                backbone=_backbone(depth), neck=_neck(depth), bbox_head=head,
                train_cfg=copy.deepcopy(_TRAIN), test_cfg=copy.deepcopy(_TEST))


def ldv2_model(student='r50', teacher='x101dcn'):
    """BASELINE.json configs[3]: "LDv2 R50 / X101-DCN teacher with VLR + IMLoss(finegrain)" -- no such file ships with
    the reference; composed as SURVEY.md 8d prescribes from configs/ldv2/ld_r50_gflv2_r101_fpn_1x.py (LDv2Head student:
    reg_topk 4, reg_channels 64, add_mean, QFL on joint scores) with configs/imv2/gflv2_x101_fpn_2x_coco.py as teacher
    (GFocalHead on ResNeXt-101 32x4d, DCN in stages 3-4), loss_ld_vlr + IMLoss(2.0), imitation_method='finegrained'."""
    m = ld_model(student)
    v2 = dict(reg_topk=4, reg_channels=64, add_mean=True,
              loss_cls=dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0))
    m['bbox_head'].update(type='LDv2Head', imitation_method='finegrained', loss_im=dict(type='IMLoss', loss_weight=2.0), **v2)
    t = m['teacher_config']['model']
    t['bbox_head'].update(type='GFocalHead', **copy.deepcopy(v2))
    if teacher == 'x101dcn':
        t['pretrained'] = 'open-mmlab://resnext101_32x4d'
        t['backbone'] = dict(type='ResNeXt', depth=101, groups=32, base_width=4, num_stages=4, out_indices=(0, 1, 2, 3),
                             frozen_stages=1, norm_cfg=dict(type='BN', requires_grad=True),
                             dcn=dict(type='DCN', deform_groups=1, fallback_on_stride=False),
                             stage_with_dcn=(False, False, True, True), norm_eval=True, style='pytorch')
    elif teacher != 'r101':
        raise KeyError(teacher)
    return m

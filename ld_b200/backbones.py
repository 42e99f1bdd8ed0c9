"""ResNet backbone and FPN neck registered under the reference's names, constructor kwargs and
parameter names (checkpoint keys `backbone.layer2.0.conv1.weight`, `neck.lateral_convs.0.conv.weight`,
...), executing on the tcgen05 convolution kernels.

Reference: mmdet/models/backbones/resnet.py:303-649, mmdet/models/utils/res_layer.py:24-103,
mmdet/models/necks/fpn.py:66-221.  Tensors cross module boundaries as NCHW-*shaped* bf16 tensors with
channels-last strides (zero-copy views of the NHWC buffers the kernels work on).
"""
import math

import torch
from torch import nn

from . import engine, ops
from .registry import BACKBONES, NECKS


def nhwc(x):
    """NCHW-shaped tensor -> contiguous NHWC in the activation dtype of the current precision mode (bf16, or fp32 in
    the parity mode); zero-copy for channels-last inputs of that dtype."""
    y = x.permute(0, 2, 3, 1)
    dt = engine.act_dtype()
    if y.dtype != dt:
        y = y.to(dt)
    return y if y.is_contiguous() else y.contiguous()


def nchw_view(y):
    return y.permute(0, 3, 1, 2)


class _Conv2d(nn.Conv2d):
    """nn.Conv2d parameter container (same state_dict keys / init) whose weight lives channels-last."""

    def _cl(self):
        if not self.weight.is_contiguous(memory_format=torch.channels_last):
            engine.as_channels_last_param(self.weight)
        return self.weight


class _Block(nn.Module):
    def _bn_affine(self, bn):
        return engine.bn_affine(bn)


class BasicBlock(_Block):
    """resnet.py:13-92."""
    expansion = 1

    def __init__(self, inplanes, planes, stride=1, downsample=None):
        super().__init__()
        self.conv1 = _Conv2d(inplanes, planes, 3, stride, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(planes)
        self.conv2 = _Conv2d(planes, planes, 3, 1, 1, bias=False)
        self.bn2 = nn.BatchNorm2d(planes)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        ds = None if self.downsample is None else (self.downsample[0], self.downsample[1])
        return engine.res_block(x, [(self.conv1, self.bn1, self.stride, 1), (self.conv2, self.bn2, 1, 1)], ds)


class DeformConv2dPack(nn.Module):
    """mmcv.ops.DeformConv2dPack (type 'DCN'; mmcv is not vendored in the reference -- restated from mmcv 1.2.x):
    parameters `weight` [Cout, Cin/groups, k, k] (no bias) and `conv_offset.{weight,bias}` (a plain conv with
    deform_groups*2*k*k output channels, zero-initialised).  Forward only (the DCN backbones of the LD configs are frozen
    teachers): offset conv (tcgen05) -> bilinear-sampled im2col (ldb200_deform_im2col) -> tcgen05 GEMM."""

    def __init__(self, in_channels, out_channels, kernel_size, stride=1, padding=0, dilation=1, groups=1,
                 deform_groups=1, bias=False):
        super().__init__()
        assert not bias and in_channels % groups == 0 and out_channels % groups == 0
        self.in_channels, self.out_channels, self.kernel_size = in_channels, out_channels, (kernel_size, kernel_size)
        self.stride, self.padding, self.dilation = stride, padding, dilation
        self.groups, self.deform_groups = groups, deform_groups
        self.weight = nn.Parameter(torch.empty(out_channels, in_channels // groups, kernel_size, kernel_size))
        n = in_channels * kernel_size * kernel_size
        nn.init.uniform_(self.weight, -1.0 / math.sqrt(n), 1.0 / math.sqrt(n))      # DeformConv2d.reset_parameters
        self.conv_offset = _Conv2d(in_channels, deform_groups * 2 * kernel_size * kernel_size, kernel_size, stride,
                                   padding, dilation=dilation, bias=True)
        nn.init.zeros_(self.conv_offset.weight)
        nn.init.zeros_(self.conv_offset.bias)

    def forward_nhwc(self, x, scale=None, shift=None, relu=False):
        """x [N,H,W,Cin] NHWC (bf16, or fp32 in the parity mode) -> relu?(scale * dcn(x) + shift) [N,HO,WO,Cout]."""
        if any(p.requires_grad for p in self.parameters()) and torch.is_grad_enabled():
            raise NotImplementedError('DeformConv2dPack runs forward-only (the DCN backbones of the LD configs are frozen '
                                      'teachers, kd_one_stage.py:70-72)')
        k = self.kernel_size[0]
        n, h, w, c = x.shape
        co_off = self.conv_offset.out_channels
        off = ops.conv2d_tc(x, engine.packed_weight(self.conv_offset._cl()), co_off, k, k, self.stride, self.padding,
                            out_dtype=torch.float32, bias=self.conv_offset.bias.detach().float().contiguous())
        ho, wo = off.shape[1:3]
        col = ops.deform_im2col(x, off, k, k, self.stride, self.padding, self.dilation, self.deform_groups)
        wp = engine.packed_weight(self.weight, groups=self.groups)
        y = ops.conv2d_tc(col.view(1, 1, n * ho * wo, k * k * c), wp, self.out_channels, scale=scale, bias=shift, relu=relu)
        return y.view(n, ho, wo, self.out_channels)


class Bottleneck(_Block):
    """resnet.py:95-300, style='pytorch' (the stride sits on the 3x3 conv); with `groups` / `base_width` the ResNeXt
    bottleneck (resnext.py:11-106), with `dcn` its 3x3 conv is a DeformConv2dPack (resnet.py:172-184)."""
    expansion = 4

    def __init__(self, inplanes, planes, stride=1, downsample=None, groups=1, base_width=4, base_channels=64, dcn=None):
        super().__init__()
        width = planes if groups == 1 else math.floor(planes * (base_width / base_channels)) * groups
        self.width, self.groups = width, groups
        self.with_dcn = dcn is not None
        self.conv1 = _Conv2d(inplanes, width, 1, bias=False)
        self.bn1 = nn.BatchNorm2d(width)
        if self.with_dcn:
            dcn = dict(dcn)
            if dcn.pop('type', 'DCN') != 'DCN':
                raise NotImplementedError('only dcn=dict(type="DCN") (DCNv1) is used by the LD / GFL configs')
            if dcn.pop('fallback_on_stride', False):
                raise NotImplementedError('dcn.fallback_on_stride')
            self.conv2 = DeformConv2dPack(width, width, 3, stride, 1, groups=groups, **dcn)
        else:
            self.conv2 = _Conv2d(width, width, 3, stride, 1, groups=groups, bias=False)
        self.bn2 = nn.BatchNorm2d(width)
        self.conv3 = _Conv2d(width, planes * 4, 1, bias=False)
        self.bn3 = nn.BatchNorm2d(planes * 4)
        self.downsample = downsample
        self.stride = stride

    def forward(self, x):
        if self.groups > 1 or self.with_dcn:
            return self._forward_frozen(x)
        ds = None if self.downsample is None else (self.downsample[0], self.downsample[1])
        return engine.res_block(x, [(self.conv1, self.bn1, 1, 0), (self.conv2, self.bn2, self.stride, 1),
                                    (self.conv3, self.bn3, 1, 0)], ds)

    def _forward_frozen(self, x):
        """Grouped / deformable 3x3: forward-only path (the X101-32x4d-DCN and R101-DCN backbones are frozen teachers).
        Grouped weights run as block-diagonal dense tcgen05 tiles."""
        if torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError('grouped / deformable bottlenecks are implemented forward-only (frozen teacher)')

        def aff(bn):
            s_, b_ = engine.bn_affine(bn)
            return s_.detach().contiguous(), b_.detach().contiguous()
        s1, b1 = aff(self.bn1)
        s2, b2 = aff(self.bn2)
        s3, b3 = aff(self.bn3)
        idt = x
        if self.downsample is not None:
            sd, bd = aff(self.downsample[1])
            wd = self.downsample[0]
            idt = ops.conv2d_tc(x, engine.packed_weight(wd._cl()), wd.out_channels, 1, 1, self.stride, 0, scale=sd, bias=bd)
        h = ops.conv2d_tc(x, engine.packed_weight(self.conv1._cl()), self.width, scale=s1, bias=b1, relu=True)
        if self.with_dcn:
            h = self.conv2.forward_nhwc(h, s2, b2, relu=True)
        else:
            h = ops.conv2d_tc(h, engine.packed_weight(self.conv2._cl(), groups=self.groups), self.width, 3, 3, self.stride, 1,
                              scale=s2, bias=b2, relu=True)
        return ops.conv2d_tc(h, engine.packed_weight(self.conv3._cl()), self.conv3.out_channels, scale=s3, bias=b3,
                             add=idt, relu=True)


@BACKBONES.register_module()
class ResNet(nn.Module):
    arch_settings = {18: (BasicBlock, (2, 2, 2, 2)), 34: (BasicBlock, (3, 4, 6, 3)),
                     50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)),
                     152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, depth, in_channels=3, stem_channels=64, base_channels=64, num_stages=4,
                 strides=(1, 2, 2, 2), dilations=(1, 1, 1, 1), out_indices=(0, 1, 2, 3), style='pytorch',
                 deep_stem=False, avg_down=False, frozen_stages=-1, conv_cfg=None,
                 norm_cfg=dict(type='BN', requires_grad=True), norm_eval=True, dcn=None,
                 stage_with_dcn=(False, False, False, False), plugins=None, with_cp=False,
                 zero_init_residual=True):
        super().__init__()
        if depth not in self.arch_settings:
            raise KeyError(f'invalid depth {depth} for resnet')
        unsupported = dict(in_channels=in_channels != 3, stem_channels=stem_channels != 64,
                           base_channels=base_channels != 64, strides=tuple(strides) != (1, 2, 2, 2)[:num_stages],
                           dilations=any(d != 1 for d in dilations), style=style != 'pytorch', deep_stem=deep_stem,
                           avg_down=avg_down, conv_cfg=conv_cfg is not None, norm=norm_cfg.get('type') != 'BN',
                           norm_eval=not norm_eval, plugins=plugins is not None, with_cp=with_cp)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f'ld_b200 ResNet implements the LD/GFL configuration only; unsupported: {bad}')
        assert 1 <= num_stages <= 4 and max(out_indices) < num_stages
        self.depth, self.num_stages, self.out_indices = depth, num_stages, tuple(out_indices)
        self.frozen_stages, self.norm_eval, self.zero_init_residual = frozen_stages, norm_eval, zero_init_residual
        block, stage_blocks = self.arch_settings[depth]
        self.block = block
        self.dcn, self.stage_with_dcn = dcn, tuple(stage_with_dcn)
        if dcn is not None:
            assert len(self.stage_with_dcn) == num_stages and block is Bottleneck, 'DCN needs Bottleneck stages'
        self.conv1 = _Conv2d(3, 64, 7, 2, 3, bias=False)
        self.bn1 = nn.BatchNorm2d(64)
        inplanes = 64
        self.res_layers = []
        for i, nb in enumerate(stage_blocks[:num_stages]):
            planes = 64 * 2 ** i
            stride = strides[i]
            blocks = []
            for j in range(nb):
                st = stride if j == 0 else 1
                ds = None
                if j == 0 and (st != 1 or inplanes != planes * block.expansion):
                    ds = nn.Sequential(_Conv2d(inplanes, planes * block.expansion, 1, st, bias=False),
                                       nn.BatchNorm2d(planes * block.expansion))
                blocks.append(self.make_block(block, inplanes, planes, st, ds, i))
                inplanes = planes * block.expansion
            name = f'layer{i + 1}'
            self.add_module(name, nn.Sequential(*blocks))
            self.res_layers.append(name)
        self.feat_dim = inplanes
        if not norm_cfg.get('requires_grad', True):
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    for p in m.parameters():
                        p.requires_grad = False
        self._freeze_stages()

    def make_block(self, block, inplanes, planes, stride, downsample, stage):
        kw = {}
        if self.dcn is not None and self.stage_with_dcn[stage]:
            kw['dcn'] = self.dcn
        return block(inplanes, planes, stride, downsample, **kw)

    def _freeze_stages(self):
        """resnet.py:572-588."""
        if self.frozen_stages >= 0:
            for m in (self.conv1, self.bn1):
                m.eval()
                for p in m.parameters():
                    p.requires_grad = False
        for i in range(1, self.frozen_stages + 1):
            m = getattr(self, f'layer{i}')
            m.eval()
            for p in m.parameters():
                p.requires_grad = False

    def init_weights(self, pretrained=None):
        """resnet.py:590-620 (pretrained checkpoints are loaded by the caller via load_state_dict)."""
        if isinstance(pretrained, str):
            raise NotImplementedError('no network on the B200 box: load weights with load_state_dict / seeded_init')
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, nn.BatchNorm2d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        for m in self.modules():      # resnet.py:604-608: the offset branch of every DCN starts at zero
            if isinstance(m, DeformConv2dPack):
                nn.init.zeros_(m.conv_offset.weight)
                nn.init.zeros_(m.conv_offset.bias)
        if self.zero_init_residual:
            for m in self.modules():
                if isinstance(m, Bottleneck):
                    nn.init.constant_(m.bn3.weight, 0)
                elif isinstance(m, BasicBlock):
                    nn.init.constant_(m.bn2.weight, 0)

    def train(self, mode=True):
        """resnet.py:639-649: BN layers stay in eval mode (norm_eval=True)."""
        super().train(mode)
        self._freeze_stages()
        if mode and self.norm_eval:
            for m in self.modules():
                if isinstance(m, nn.BatchNorm2d):
                    m.eval()
        return self

    def stem(self, img=None, im2col=None):
        """conv1 7x7/2 + BN + ReLU + 3x3/2 max-pool (resnet.py:626-630) -> NHWC bf16."""
        if im2col is None:
            im2col = ops.stem_im2col(img.float().contiguous(), engine.act_dtype())
        a, ho, wo = im2col
        n = a.shape[0] // (ho * wo)
        w = self.conv1.weight.detach()
        wp = torch.zeros(64, 192, dtype=torch.float32, device=w.device)
        wp[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147)
        wp = ops.split_weight(wp) if engine.parity() else wp.to(torch.bfloat16)
        s, b = engine.bn_affine(self.bn1)
        y = ops.conv2d_tc(a.view(1, 1, -1, 192), wp, 64, scale=s.detach().contiguous(), bias=b.detach().contiguous(),
                          relu=True).view(n, ho, wo, 64)
        return ops.maxpool3x3s2(y)

    def stem_params(self):
        """(weights [64,192] bf16 with K = (r*7+s)*3+c zero-padded from 147, (scale, shift)) for `ops.stem_tc`."""
        w = self.conv1.weight.detach()
        wp = torch.zeros(64, 192, dtype=torch.bfloat16, device=w.device)
        wp[:, :147] = w.permute(0, 2, 3, 1).reshape(64, 147).to(torch.bfloat16)
        s, b = engine.bn_affine(self.bn1)
        return wp, (s.detach(), b.detach())

    cut_stages = False     # set by LDTrainStep when it runs the backward pass in pieces (overlapped gradient exchange)

    def forward(self, x, im2col=None, stem_out=None):
        """stem_out: conv1+BN+ReLU output [N,HO,WO,64] bf16 already computed by the fused stem kernel (opt-in)."""
        if any(p.requires_grad for p in (self.conv1.weight, self.bn1.weight)):
            raise NotImplementedError('the stem runs forward-only (frozen_stages >= 0 in every LD/GFL config)')
        with torch.no_grad():
            y = ops.maxpool3x3s2(stem_out) if stem_out is not None else self.stem(x, im2col)
        outs, cuts = [], []
        for i, name in enumerate(self.res_layers):
            y = getattr(self, name)(y)
            if i in self.out_indices:
                outs.append(nchw_view(y))
            if self.cut_stages and y.requires_grad and i + 1 < len(self.res_layers):
                # the next stage reads an ALIAS of this stage's output: a node whose backward does nothing, i.e. a place
                # where LDTrainStep can cut the backward pass (`torch.autograd.backward(inputs=[t])` also runs the
                # node that produced t)
                y = y.view_as(y)
                cuts.append(y)
        self._cuts = cuts          # inputs of the trainable stages after the first, in forward order
        return tuple(outs)


@BACKBONES.register_module()
class ResNeXt(ResNet):
    """resnext.py:109-153: ResNet with grouped 3x3 convs (groups x base_width); the teacher of BASELINE configs[3]
    (configs/imv2/gflv2_x101_fpn_2x_coco.py: depth 101, groups 32, base_width 4, DCN in stages 3-4)."""
    arch_settings = {50: (Bottleneck, (3, 4, 6, 3)), 101: (Bottleneck, (3, 4, 23, 3)), 152: (Bottleneck, (3, 8, 36, 3))}

    def __init__(self, groups=1, base_width=4, **kwargs):
        self.groups, self.base_width = groups, base_width
        super().__init__(**kwargs)

    def make_block(self, block, inplanes, planes, stride, downsample, stage):
        kw = dict(groups=self.groups, base_width=self.base_width)
        if self.dcn is not None and self.stage_with_dcn[stage]:
            kw['dcn'] = self.dcn
        return block(inplanes, planes, stride, downsample, **kw)


class _ConvOnly(nn.Module):
    """mmcv ConvModule with norm_cfg=None, act_cfg=None: just `.conv` (fpn.py:121-137)."""

    def __init__(self, cin, cout, k, stride=1, padding=0):
        super().__init__()
        self.conv = _Conv2d(cin, cout, k, stride, padding, bias=True)


@NECKS.register_module()
class FPN(nn.Module):
    def __init__(self, in_channels, out_channels, num_outs, start_level=0, end_level=-1, add_extra_convs=False,
                 extra_convs_on_inputs=True, relu_before_extra_convs=False, no_norm_on_lateral=False, conv_cfg=None,
                 norm_cfg=None, act_cfg=None, upsample_cfg=dict(mode='nearest')):
        super().__init__()
        assert isinstance(in_channels, (list, tuple))
        self.in_channels, self.out_channels, self.num_ins, self.num_outs = list(in_channels), out_channels, \
            len(in_channels), num_outs
        self.backbone_end_level = self.num_ins if end_level == -1 else end_level
        self.start_level = start_level
        if isinstance(add_extra_convs, bool) and add_extra_convs:
            add_extra_convs = 'on_input' if extra_convs_on_inputs else 'on_output'
        self.add_extra_convs = add_extra_convs
        bad = []
        if end_level != -1: bad.append('end_level')
        if add_extra_convs not in ('on_output',): bad.append(f'add_extra_convs={add_extra_convs}')
        if relu_before_extra_convs: bad.append('relu_before_extra_convs')
        if conv_cfg is not None or norm_cfg is not None or act_cfg is not None: bad.append('conv/norm/act cfg')
        if upsample_cfg.get('mode', 'nearest') != 'nearest': bad.append('upsample mode')
        if bad:
            raise NotImplementedError(f'ld_b200 FPN implements the LD/GFL configuration only; unsupported: {bad}')
        # the backbone's stage outputs are ReLU outputs of ResBlockFn nodes, which expect pre-masked gradients
        self.inputs_are_relu = True
        self.lateral_convs = nn.ModuleList()
        self.fpn_convs = nn.ModuleList()
        for i in range(self.start_level, self.backbone_end_level):
            self.lateral_convs.append(_ConvOnly(in_channels[i], out_channels, 1))
            self.fpn_convs.append(_ConvOnly(out_channels, out_channels, 3, padding=1))
        extra = num_outs - self.backbone_end_level + self.start_level
        for _ in range(extra):
            self.fpn_convs.append(_ConvOnly(out_channels, out_channels, 3, stride=2, padding=1))

    def init_weights(self):
        """fpn.py:163-167."""
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.xavier_uniform_(m.weight)
                nn.init.constant_(m.bias, 0)

    def forward(self, inputs):
        assert len(inputs) == len(self.in_channels)
        feats = [nhwc(x) for x in inputs[self.start_level:]]
        nb = len(feats)
        lat = [None] * nb
        for i in range(nb - 1, -1, -1):   # top-down: lateral conv with the coarser level's upsample fused in
            c = self.lateral_convs[i].conv
            lat[i] = engine.conv_affine_act(feats[i], c._cl(), None, c.bias, lat[i + 1] if i + 1 < nb else None,
                                            1, 0, False, x_premask=self.inputs_are_relu)
        # one row-space buffer for all output levels (ld_b200.h "row space")
        n = feats[0].shape[0]
        sizes = [tuple(l.shape[1:3]) for l in lat]
        for _ in range(len(self.fpn_convs) - nb):
            h, w = sizes[-1]
            sizes.append(((h - 1) // 2 + 1, (w - 1) // 2 + 1))
        rows = torch.empty(n * sum(h * w for h, w in sizes), self.out_channels, dtype=feats[0].dtype,
                           device=feats[0].device)
        outs, o = [], 0
        for i, (h, w) in enumerate(sizes):
            view = rows[o:o + n * h * w].view(n, h, w, self.out_channels)
            o += n * h * w
            c = self.fpn_convs[i].conv
            src = lat[i] if i < nb else outs[-1]
            outs.append(engine.conv_affine_act(src, c._cl(), None, c.bias, None, c.stride[0], 1, False, out=view))
        return tuple(nchw_view(y) for y in outs)


def rows_of(levels):
    """Per-level NCHW-shaped views laid out back to back in one storage -> the [R, C] row-space tensor
    (zero-copy); anything else is copied into a fresh row-space buffer."""
    c = levels[0].shape[1]
    try:
        base = levels[0].permute(0, 2, 3, 1)
        ptr = base.data_ptr()
        es = base.element_size()
        total = 0
        ok = True
        for t in levels:
            v = t.permute(0, 2, 3, 1)
            ok = ok and v.is_contiguous() and v.data_ptr() == ptr + total * c * es and v.dtype == base.dtype \
                and t.shape[1] == c and v.untyped_storage().data_ptr() == base.untyped_storage().data_ptr()
            total += v.shape[0] * v.shape[1] * v.shape[2]
        if ok:
            return torch.as_strided(base, (total, c), (c, 1), base.storage_offset())
    except Exception:
        pass
    return ops.to_rows(list(levels))

"""GPU: ResNeXt (grouped 3x3 convs as block-diagonal tcgen05 tiles, resnext.py:64-75) and DCNv1 (offset conv ->
ldb200_deform_im2col -> tcgen05 GEMM; mmcv DeformConv2dPack, resnet.py:172-184) -- the frozen teacher backbones of
BASELINE configs[3] (X101-32x4d-DCN) and of configs/ld/ld_r101_gflv1_r101dcn_fpn_coco_2x.py -- against the CPU oracle
(pinned to the unmodified reference + the torchvision-based mmcv stand-in by tests/test_config_dropin.py)."""
import pytest
import torch

from ld_b200 import synth

pytestmark = pytest.mark.gpu


def test_deform_im2col_matches_torchvision():
    from torchvision.ops import deform_conv2d
    from ld_b200 import ops
    g = torch.Generator().manual_seed(2)
    for (n, h, w, c, stride, dg) in ((2, 13, 17, 64, 1, 1), (1, 20, 15, 128, 2, 1), (2, 9, 11, 64, 1, 2)):
        x = torch.randn(n, c, h, w, generator=g)
        ho, wo = (h + 2 - 3) // stride + 1, (w + 2 - 3) // stride + 1
        off = torch.randn(n, 2 * dg * 9, ho, wo, generator=g) * 2.5          # samples leave the image at the borders
        wt = torch.randn(32, c, 3, 3, generator=g) * 0.1
        want = deform_conv2d(x, off, wt, None, stride, 1, 1)                  # [n, 32, ho, wo]
        col = ops.deform_im2col(x.permute(0, 2, 3, 1).contiguous().cuda(), off.permute(0, 2, 3, 1).contiguous().cuda(),
                                3, 3, stride, 1, 1, dg)
        got = col.float() @ wt.permute(0, 2, 3, 1).reshape(32, -1).t().cuda()
        got = got.view(n, ho, wo, 32).permute(0, 3, 1, 2).cpu()
        assert float((got - want).abs().max()) <= 1e-4 * float(want.abs().max()), (n, h, w, c, stride, dg)


@pytest.mark.parametrize('mode,tol', [('bf16', 4e-2), ('fp32', 1e-4)])
@pytest.mark.parametrize('kind', ['x50_32x4d_dcn', 'r50_dcn'])
def test_grouped_dcn_backbone_matches_oracle(kind, mode, tol):
    import ld_b200
    from ld_b200 import engine
    from ld_b200.registry import build_backbone
    from oracle import ld_oracle as lo
    common = dict(depth=50, num_stages=4, out_indices=(0, 1, 2, 3), frozen_stages=4, norm_eval=True, style='pytorch',
                  norm_cfg=dict(type='BN', requires_grad=False), dcn=dict(type='DCN', deform_groups=1, fallback_on_stride=False))
    if kind == 'x50_32x4d_dcn':
        m = build_backbone(dict(type='ResNeXt', groups=32, base_width=4, stage_with_dcn=(False, False, True, True), **common))
        om = lo.OracleResNet(50, groups=32, base_width=4, stage_with_dcn=(False, False, True, True))
    else:
        m = build_backbone(dict(type='ResNet', stage_with_dcn=(False, True, True, True), **common))
        om = lo.OracleResNet(50, stage_with_dcn=(False, True, True, True))
    synth.seeded_init(m, 13)
    synth.seeded_init(om, 13)
    m, om = m.cuda().eval(), om.eval()
    x = torch.randn(2, 3, 128, 160, generator=torch.Generator().manual_seed(4))
    with torch.no_grad(), engine.precision(mode):
        got = m(x.cuda())
    with torch.no_grad():
        want = om(x)
    torch.cuda.synchronize()
    for lvl, (a, b) in enumerate(zip(got, want)):
        err = float((a.float().cpu() - b).norm() / b.norm())
        assert err <= tol, (kind, mode, lvl, err)

"""GPU: the loss modules of the registry surface called on their own -- `build_loss(cfg)(pred, target, weight,
avg_factor, reduction_override)` -- against golden vectors of the UNMODIFIED reference modules
(gfocal_loss.py:77-179, iou_loss.py:325-360; tests/golden/m_*.pt from oracle/make_golden.py --loss-modules):
values and d/d pred for every reduction variant, fp32, 1e-4."""
import os

import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', cases.LOSS_MODULE_CASES, ids=lambda c: c[0])
def test_standalone_loss_modules_match_reference(golden_dir, case):
    from ld_b200.registry import build_loss
    name, rows, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    d = {k: v.cuda() for k, v in cases.loss_module_inputs(rows, seed).items()}
    mods = dict(qfl=build_loss(dict(type='QualityFocalLoss', use_sigmoid=True, beta=2.0, loss_weight=1.0)),
                qflp=build_loss(dict(type='QualityFocalLoss', use_sigmoid=False, beta=2.0, loss_weight=1.0)),
                dfl=build_loss(dict(type='DistributionFocalLoss', loss_weight=0.25)),
                giou=build_loss(dict(type='GIoULoss', loss_weight=2.0)))
    for tag, kw in cases.LOSS_MODULE_CALLS:
        kw = dict(kw)
        w = None if kw.pop('no_weight', False) else d['weight']
        for key, pred, target in (('qfl', d['qfl_pred'], (d['label'], d['score'])),
                                  ('qflp', d['qfl_prob'], (d['label'], d['score'])),
                                  ('dfl', d['dfl_pred'], d['dfl_target']),
                                  ('giou', d['box_pred'], d['box_target'])):
            p = pred.clone().requires_grad_(True)
            out = mods[key](p, target, weight=w, **kw)
            up = torch.linspace(0.5, 1.5, out.numel()).reshape(out.shape).cuda() if out.dim() else torch.tensor(1.0).cuda()
            (out * up).sum().backward()
            torch.cuda.synchronize()
            g = gold[f'{key}_{tag}']
            assert out.shape == g['out'].shape, (key, tag)
            assert torch.allclose(out.detach().cpu(), g['out'], rtol=1e-4, atol=1e-6), (key, tag, out, g['out'])
            err = float((p.grad.cpu() - g['grad']).abs().max())
            assert err <= 1e-4 * float(g['grad'].abs().max()) + 1e-9, (key, tag, err)


def test_avg_factor_with_sum_is_an_error():
    from ld_b200.registry import build_loss
    m = build_loss(dict(type='DistributionFocalLoss', loss_weight=0.25))
    with pytest.raises(ValueError):      # losses/utils.py:52-53
        m(torch.randn(8, 17).cuda(), torch.rand(8).cuda() * 15, avg_factor=4.0, reduction_override='sum')

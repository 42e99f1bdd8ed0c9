"""GPU: the inference step after the training path -- `GFLHead.get_bboxes` (ldb200_gfl_get_bboxes: exact
per-level top-k, Integral, distance2bbox + clip, score threshold, per-class NMS on mmcv's class-offset boxes) against the
golden vectors of the unmodified reference head (tests/golden/i_*.pt) and `simple_test` end to end."""
import os

import pytest
import torch

from oracle import cases

pytestmark = pytest.mark.gpu


def make_head():
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    cfg = ConfigDict(presets.ld_model('r50'))
    hc = cfg.bbox_head
    hc['train_cfg'], hc['test_cfg'] = cfg.train_cfg, cfg.test_cfg
    return ld_b200.build_head(hc).cuda().eval()


@pytest.mark.parametrize('case', cases.INFER_CASES, ids=lambda c: c[0])
def test_get_bboxes_matches_reference(golden_dir, case):
    name, n, h, w, seed, rescale = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    d = cases.infer_case_inputs(n, h, w, seed)
    head = make_head()
    res = head.get_bboxes([t.cuda() for t in d['cls']], [t.cuda() for t in d['reg']], d['img_metas'], rescale=rescale)
    torch.cuda.synchronize()
    for (det, lab), gd, gl in zip(res, gold['dets'], gold['labels']):
        assert det.shape == gd.shape and torch.equal(lab.cpu(), gl)           # which (anchor, class) pairs survive: exact
        assert torch.allclose(det[:, 4].cpu(), gd[:, 4], rtol=1e-6, atol=1e-7)   # scores
        assert torch.allclose(det[:, :4].cpu(), gd[:, :4], rtol=1e-5, atol=2e-4)  # boxes (pixels)


def test_select_handles_small_levels_and_empty_results():
    """Levels with fewer anchors than nms_pre are taken whole; an image without any score above the threshold returns
    zero detections."""
    head = make_head()
    d = cases.infer_case_inputs(2, 128, 160, 5)
    cls = [t.cuda() for t in d['cls']]
    cls = [c.clone() for c in cls]
    for c in cls:
        c[1] = -20.0                       # image 1: nothing passes score_thr
    res = head.get_bboxes(cls, [t.cuda() for t in d['reg']], d['img_metas'])
    assert res[1][0].shape == (0, 5) and res[1][1].numel() == 0 and res[0][0].shape[0] > 0


def test_simple_test_end_to_end():
    import ld_b200
    from ld_b200 import presets, synth
    from ld_b200.config import ConfigDict
    m = ld_b200.build_detector(ConfigDict(presets.gfl_model(18)))
    synth.seeded_init(m, 3)
    m = m.cuda().eval()
    b = synth.make_batch(2, 128, 160, seed=8)
    with torch.no_grad():     # random weights: shift the class bias so that ~0.5 % of the scores pass score_thr = 0.05
        logits = torch.cat([c.flatten() for c in m.bbox_head(m.extract_feat(b['img'].cuda()))[0]])
        q = float(torch.quantile(logits[:2000000].float(), 0.995))
        m.bbox_head.gfl_cls.bias += (-2.94 - q)
    out = m(b['img'].cuda(), b['img_metas'], return_loss=False)
    assert len(out) == 2 and len(out[0]) == 80 and all(a.shape[1] == 5 for a in out[0])
    assert 0 < sum(a.shape[0] for a in out[0]) <= 100

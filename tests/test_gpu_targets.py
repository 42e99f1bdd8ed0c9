"""GPU: ldb200_ld_targets (C-ABI) vs golden vectors from the unmodified reference (bit-exact) and vs
the CPU oracle on further seeds, edge cases and the full benchmark size."""
import os

import pytest
import torch

from ld_b200 import synth
from oracle import cases, ld_oracle as lo

pytestmark = pytest.mark.gpu


def run_targets(n, h, w, gb, gl, im_mode=0, pad_hw=None):
    from ld_b200 import lib as L, ops
    sizes = cases.featmap_sizes(h, w)
    pyr = L.Pyramid.make(n, sizes, cases.STRIDES)
    pad_hw = pad_hw or [(h, w)] * n
    t = ops.ld_targets(pyr, gb, gl, pad_hw, topk=9, im_mode=im_mode, num_classes=80,
                       device=torch.device('cuda'))
    torch.cuda.synchronize()
    return pyr, t


def cat_levels(v, pyr):
    from ld_b200 import ops
    return torch.cat(ops.rows_to_levels(v, pyr), 1).cpu()


def compare_with_oracle(n, h, w, gb, gl, pad_hw=None):
    pyr, t = run_targets(n, h, w, gb, gl, 0, pad_hw)
    _, tc = run_targets(n, h, w, gb, gl, 1, pad_hw)
    sizes = cases.featmap_sizes(h, w)
    metas = synth.make_img_metas(n, h, w)
    if pad_hw:
        for m, p in zip(metas, pad_hw):
            m['pad_shape'] = (p[0], p[1], 3)
    o = lo.ld_targets(sizes, cases.STRIDES, metas, gb, gl, imitation='finegrained')
    oc = lo.ld_targets(sizes, cases.STRIDES, metas, gb, gl, imitation='fitnet')
    assert torch.equal(cat_levels(t.labels, pyr).long(), torch.cat(o['labels'], 1))
    assert torch.equal(cat_levels(t.gt_inds, pyr).long(), torch.cat(o['gt_inds'], 1))
    assert torch.equal(cat_levels(t.label_weights, pyr), torch.cat(o['label_weights'], 1))
    assert torch.equal(cat_levels(t.bbox_targets, pyr), torch.cat(o['bbox_targets'], 1))
    assert torch.equal(cat_levels(t.vlr, pyr), torch.cat(o['vlr'], 1))
    assert torch.equal(cat_levels(t.im, pyr), torch.cat(o['im'], 1))
    assert torch.equal(cat_levels(tc.im, pyr), torch.cat(oc['im'], 1))
    cnt = t.counts.cpu()
    assert int(cnt[2]) == o['num_total_pos']
    # compact lists are exactly the pos|vlr rows and the fg rows (order free)
    lab = t.labels.cpu()
    vlr = t.vlr.cpu()
    im = t.im.cpu()
    want = ((lab < 80) | (vlr > 0)).nonzero().squeeze(1)
    got = t.sparse_rows[:int(cnt[0])].cpu().long().sort()[0]
    assert torch.equal(got, want)
    assert torch.equal(t.fg_rows[:int(cnt[1])].cpu().long().sort()[0], (im > 0).nonzero().squeeze(1))
    o_pos = [int((l < 80).sum()) for l in o['labels']]
    assert [int(cnt[4 + 3 * l]) for l in range(5)] == o_pos
    return t


@pytest.mark.parametrize('case', cases.TARGET_CASES, ids=lambda c: c[0])
def test_targets_golden_bit_exact(golden_dir, case):
    name, n, h, w, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    gb, gl = synth.make_gts(n, h, w, seed)
    pyr, t = run_targets(n, h, w, gb, gl, 0)
    _, tc = run_targets(n, h, w, gb, gl, 1)
    assert torch.equal(cat_levels(t.labels, pyr).to(torch.int16), gold['labels'])
    assert torch.equal(cat_levels(t.label_weights, pyr).to(torch.uint8), gold['label_weights'])
    assert torch.equal(cat_levels(t.bbox_targets, pyr), gold['bbox_targets'])
    assert torch.equal(cat_levels(t.vlr, pyr), gold['vlr'])
    assert torch.equal(cat_levels(t.im, pyr).to(torch.uint8), gold['im_finegrained'])
    assert torch.equal(cat_levels(tc.im, pyr).to(torch.uint8), gold['im_center'])
    assert int(t.counts[2]) == gold['num_total_pos']


@pytest.mark.parametrize('seed', list(range(100, 124)))
def test_targets_vs_oracle_seeds(seed):
    h, w = [(256, 320), (320, 448), (512, 512), (416, 640)][seed % 4]
    n = 1 + seed % 3
    gb, gl = synth.make_gts(n, h, w, seed, mean_boxes=[1.0, 7.0, 20.0][seed % 3])
    compare_with_oracle(n, h, w, gb, gl)


def test_targets_edge_cases():
    # single gt (G == 1 takes torch's inner-vectorised sum order), tiny and huge boxes, many gts
    h, w = 256, 320
    one = [torch.tensor([[40.3, 50.7, 200.9, 180.2]])]
    compare_with_oracle(1, h, w, one, [torch.tensor([3])])
    tiny = [torch.tensor([[100.2, 100.1, 104.9, 105.3], [0.0, 0.0, 319.5, 255.5]])]
    compare_with_oracle(1, h, w, tiny, [torch.tensor([1, 79])])
    gb, gl = synth.make_gts(2, 384, 384, 77, mean_boxes=60.0)
    compare_with_oracle(2, 384, 384, gb, gl)
    # pad_shape smaller than the batch tensor -> invalid anchors are filled (unmap)
    gb, gl = synth.make_gts(2, 200, 260, 78)
    compare_with_oracle(2, 256, 320, gb, gl, pad_hw=[(200, 260), (256, 320)])


def test_targets_full_size_properties():
    """BASELINE size (16 x 800x1344): oracle on 2 of the images + size-independent invariants."""
    n, h, w = 16, 800, 1344
    gb, gl = synth.make_gts(n, h, w, 5)
    pyr, t = run_targets(n, h, w, gb, gl)
    cnt = t.counts.cpu()
    lab = t.labels.cpu()
    gi = t.gt_inds.cpu()
    bt = t.bbox_targets.cpu()
    vlr = t.vlr.cpu()
    pos = lab < 80
    assert int(cnt[3]) == int(pos.sum()) and int(cnt[0]) == int((pos | (vlr > 0)).sum())
    assert torch.equal(pos, gi > 0)
    assert (bt[~pos] == 0).all() and ((vlr >= 0) & (vlr <= 1)).all()
    # every positive's target box is one of its image's gts and carries that gt's label
    levels = t.pyr.level_anchors
    o = 0
    for A in levels:
        for i in range(n):
            sl = slice(o + i * A, o + (i + 1) * A)
            p = pos[sl].nonzero().squeeze(1)
            if p.numel():
                g = gi[sl][p].long() - 1
                assert torch.equal(bt[sl][p], gb[i][g]) and torch.equal(lab[sl][p].long(), gl[i][g])
        o += n * A
    # idempotence
    _, t2 = run_targets(n, h, w, gb, gl)
    assert torch.equal(t2.labels, t.labels) and torch.equal(t2.vlr, t.vlr)
    sub = [0, 9]
    pyr2, ts = run_targets(2, h, w, [gb[i] for i in sub], [gl[i] for i in sub])
    o = lo.ld_targets(cases.featmap_sizes(h, w), cases.STRIDES, synth.make_img_metas(2, h, w),
                      [gb[i] for i in sub], [gl[i] for i in sub])
    assert torch.equal(cat_levels(ts.labels, pyr2).long(), torch.cat(o['labels'], 1))
    assert torch.equal(cat_levels(ts.vlr, pyr2), torch.cat(o['vlr'], 1))

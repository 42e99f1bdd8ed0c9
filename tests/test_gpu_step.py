"""GPU: a full LD train step through the reference-shaped modules (tcgen05 convs in bf16) against the golden
vectors of the unmodified fp32 reference and the CPU oracle.  bf16 operands bound the agreement: loss terms
within 3 %, gradient direction (cosine) > 0.99, see DESIGN.md "precision"."""
import os

import pytest
import torch

from ld_b200 import synth
from oracle import cases

pytestmark = pytest.mark.gpu


def build(sd, td, seed):
    import ld_b200
    from ld_b200 import presets
    from ld_b200.config import ConfigDict
    m = ld_b200.build_detector(ConfigDict(presets.ld_model(f'r{sd}', teacher_depth=td)))
    synth.seeded_init(m, seed)
    synth.seeded_init(m.teacher_model, seed + 1, prefix='teacher.')
    return m.cuda().train()


def to_dev(batch):
    return dict(img=batch['img'].cuda(), img_metas=batch['img_metas'],
                gt_bboxes=[b.cuda() for b in batch['gt_bboxes']], gt_labels=[b.cuda() for b in batch['gt_labels']])


@pytest.mark.parametrize('case', cases.STEP_CASES, ids=lambda c: c[0])
def test_step_vs_golden(golden_dir, case):
    name, sd, td, n, h, w, seed = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    m = build(sd, td, seed)
    data = to_dev(synth.make_batch(n, h, w, seed=1000 + seed))
    losses = m(**data)
    loss, log = m._parse_losses(losses)
    loss.backward()
    torch.cuda.synchronize()
    for k, v in gold['log_vars'].items():
        assert abs(log[k] - v) <= 0.03 * abs(v) + 2e-3, (k, log[k], v)
    named = dict(m.named_parameters())
    got_none = sorted(k for k, p in named.items() if p.grad is None)
    assert got_none == sorted(gold['no_grad'])
    for k, g in gold['grads'].items():
        mine = named[k].grad.float().cpu().flatten()
        ref = g.flatten()
        cos = float(torch.dot(mine, ref) / (mine.norm() * ref.norm() + 1e-30))
        assert cos > 0.99, (k, cos)
        assert abs(float(mine.norm()) - float(ref.norm())) <= 0.05 * float(ref.norm()), k
    bad = []
    for k, st in gold['grad_stats'].items():
        gn = float(named[k].grad.double().norm())
        if abs(gn - float(st[0])) > 0.08 * float(st[0]) + 1e-7:
            bad.append((k, gn, float(st[0])))
    assert len(bad) <= len(gold['grad_stats']) // 20, bad[:10]


def test_train_steps_reduce_loss_and_match_sgd():
    """LDTrainStep (flat-buffer fused SGD) == torch.optim.SGD on the same gradients, params/grads stay
    aliased to the flat buffers, and the loss goes down on a fixed batch."""
    import ld_b200
    m = build(18, 18, 3)
    ref = build(18, 18, 3)
    data = to_dev(synth.make_batch(2, 128, 160, seed=77))
    ts = ld_b200.LDTrainStep(m, lr=0.0005, warmup_iters=0)
    opt = torch.optim.SGD([p for p in ref.parameters() if p.requires_grad], lr=0.0005, momentum=0.9, weight_decay=1e-4)
    p = m.bbox_head.gfl_cls.weight
    assert p.data_ptr() >= ts.flat_w.data_ptr() and p.grad.data_ptr() >= ts.flat_g.data_ptr()
    logs = []
    for it in range(3):
        loss, log = ts.step(data)
        logs.append(log['loss'])
        assert all(torch.isfinite(torch.tensor(list(log.values())))), log
        if it < 2:
            opt.zero_grad(set_to_none=True)
            l2, _ = ref._parse_losses(ref(**data))
            l2.backward()
            opt.step()
            for q in ref.parameters():          # torch updated the weights: refresh the bf16 operand copies
                q.__dict__.pop('_ldb_packed', None)
                q.__dict__.pop('_ldb_dgrad', None)
            for (k, a), (_, b) in zip(m.named_parameters(), ref.named_parameters()):
                if a.requires_grad:
                    err = float((a - b).abs().max())
                    assert err <= 2e-3 * float(b.abs().max()) + 1e-6, (it, k, err)
    assert min(logs[1:3]) < logs[0], logs
    assert float(ts.flat_g.abs().sum()) > 0


def test_cuda_graph_step_matches_eager():
    """LDTrainStep.capture()/step_graphed(): the whole iteration (forward, fused loss, backward, SGD) replayed
    from one CUDA graph gives the same parameters as the eager path, also when the batch (and its number of
    gt boxes) changes between replays."""
    import ld_b200
    a, b = build(18, 18, 5), build(18, 18, 5)
    batches = [to_dev(synth.make_batch(2, 128, 160, seed=s)) for s in (81, 82, 83)]
    ta = ld_b200.LDTrainStep(a, lr=0.0005, warmup_iters=2)
    tb = ld_b200.LDTrainStep(b, lr=0.0005, warmup_iters=2)
    tb.capture(batches[0], warmup=1)          # one eager iteration on batches[0] (side stream), then capture
    ta.step(batches[0], sync_logs=False)
    for i, d in enumerate(batches[1:] + batches[:1]):
        la, _ = ta.step(d, sync_logs=False)
        lb = tb.step_graphed(d)
        torch.cuda.synchronize()
        assert abs(float(la) - float(lb)) <= 5e-3 * abs(float(la)), (i, float(la), float(lb))
    for (k, p), (_, q) in zip(a.named_parameters(), b.named_parameters()):
        if p.requires_grad:
            err = float((p - q).abs().max())
            assert err <= 2e-3 * float(p.abs().max()) + 1e-6, (k, err)

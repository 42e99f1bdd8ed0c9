"""GPU: dense LD-KL fwd+bwd micro-benchmark kernel (BASELINE configs[4]) vs the oracle's
knowledge_distillation_kl_div_loss restatement + autograd."""
import pytest
import torch

from oracle import ld_oracle as lo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('rows,bins,T', [(1000, 17, 10.0), (256 * 7, 17, 10.0), (256 * 300 + 13, 17, 10.0),
                                         (0, 17, 10.0), (5, 17, 2.0), (777, 80, 2.0), (333, 8, 1.0)])
def test_kl_dense(rows, bins, T):
    from ld_b200 import ops
    g = torch.Generator().manual_seed(rows + bins)
    s = (torch.randn(rows, bins, generator=g) * 2).requires_grad_(True)
    t = torch.randn(rows, bins, generator=g) * 2
    w = torch.rand(rows, generator=g)
    ref = 0.25 * (lo.kd_kl_rows(s, t, T) * w).sum() / 4.0
    ref.backward()
    loss, ds = ops.kl_dense_fwd_bwd(s.detach().cuda(), t.cuda(), w.cuda(), T=T, loss_weight=0.25, avg_factor=4.0)
    torch.cuda.synchronize()
    assert abs(float(loss) - float(ref)) <= 1e-4 * max(abs(float(ref)), 1e-6)
    if rows:
        err = float((ds.cpu() - s.grad).abs().max())
        assert err <= 1e-4 * float(s.grad.abs().max()), err


def test_kl_dense_full_size_linearity():
    """configs[4]: 5 levels x 200k anchors x 4 sides: weights scale the loss and gradient linearly."""
    from ld_b200 import ops
    dev = torch.device('cuda')
    g = torch.Generator(device='cuda').manual_seed(0)
    rows = 5 * 200000 * 4
    s = torch.randn(rows, 17, device=dev, generator=g) * 2
    t = torch.randn(rows, 17, device=dev, generator=g) * 2
    w = torch.rand(rows, device=dev, generator=g)
    l1, d1 = ops.kl_dense_fwd_bwd(s, t, w)
    l2, d2 = ops.kl_dense_fwd_bwd(s, t, 3 * w)
    l0, d0 = ops.kl_dense_fwd_bwd(s, s, w)            # KL(p||p) = 0, zero gradient
    torch.cuda.synchronize()
    assert torch.allclose(3 * l1, l2, rtol=1e-4) and torch.allclose(3 * d1, d2, rtol=1e-4, atol=1e-9)
    assert abs(float(l0)) <= 1e-3 and float(d0.abs().max()) <= 1e-7
    # gradient rows sum to zero (softmax difference)
    assert float(d1.sum(1).abs().max()) <= 1e-6

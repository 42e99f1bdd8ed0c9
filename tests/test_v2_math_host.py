"""CPU: the row-level arithmetic of the GFLv2 / LDv2 kernels (ld_b200/csrc/v2_math.cuh, shared host+device inline
functions) compiled for the host and checked against torch autograd of the reference formulas
(gfocal_head.py:198-215 DGQP branch, gfocal_loss.py:29-49 QFL with use_sigmoid=False)."""
import ctypes as C
import os
import subprocess

import pytest
import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def shim(tmp_path_factory):
    out = str(tmp_path_factory.mktemp('shim') / 'libv2_math_host.so')
    src = os.path.join(ROOT, 'tests', 'host_shim', 'v2_math_host.cpp')
    subprocess.run(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-ffp-contract=off', src, '-o', out], check=True)
    return C.CDLL(out)


def _p(t):
    return C.c_void_p(t.data_ptr())


def dgqp_ref(reg, W1, b1, W2, b2, bins=17, topk=4):
    """gfocal_head.py:198-213 on rows: reg [R, 4*bins] -> quality [R]."""
    prob = F.softmax(reg.reshape(-1, 4, bins), dim=2)
    top = prob.topk(topk, dim=2)[0]
    stat = torch.cat([top, top.mean(dim=2, keepdim=True)], dim=2).reshape(reg.shape[0], -1)
    h = F.relu(F.linear(stat, W1, b1))
    return torch.sigmoid(F.linear(h, W2[None], b2)).squeeze(1)


def test_dgqp_forward_backward(shim):
    g = torch.Generator().manual_seed(3)
    R, bins, topk, hid = 500, 17, 4, 64
    ns = 4 * (topk + 1)
    reg = (torch.randn(R, 4 * bins, generator=g) * 2).requires_grad_(True)
    W1 = (torch.randn(hid, ns, generator=g) * 0.5).requires_grad_(True)
    b1 = (torch.randn(hid, generator=g) * 0.2).requires_grad_(True)
    W2 = (torch.randn(hid, generator=g) * 0.5).requires_grad_(True)
    b2 = torch.tensor([0.1], requires_grad=True)
    q_ref = dgqp_ref(reg, W1, b1, W2, b2, bins, topk)
    dq = torch.randn(R, generator=g)
    q_ref.backward(dq)

    q = torch.empty(R)
    shim.host_dgqp_forward(_p(reg.detach()), R, bins, topk, 1, hid, _p(W1.detach()), _p(b1.detach()), _p(W2.detach()),
                           C.c_float(float(b2)), _p(q))
    assert torch.allclose(q, q_ref.detach(), rtol=1e-5, atol=1e-6)
    d_reg, dW1, db1, dW2, db2 = (torch.empty(R, 4 * bins), torch.empty(hid, ns), torch.empty(hid), torch.empty(hid),
                                 torch.empty(1))
    shim.host_dgqp_backward(_p(reg.detach()), R, bins, topk, 1, hid, _p(W1.detach()), _p(b1.detach()),
                            _p(W2.detach()), C.c_float(float(b2)), _p(dq), _p(d_reg), _p(dW1), _p(db1), _p(dW2), _p(db2))
    for got, want, name in ((d_reg, reg.grad, 'd_reg'), (dW1, W1.grad, 'dW1'), (db1, b1.grad, 'db1'),
                            (dW2, W2.grad, 'dW2'), (db2, b2.grad, 'db2')):
        err = float((got - want).abs().max())
        assert err <= 1e-4 * float(want.abs().max()) + 1e-7, (name, err)


@pytest.mark.parametrize('beta', [2.0, 1.5])
def test_qfl_on_probabilities(shim, beta):
    g = torch.Generator().manual_seed(4)
    n = 4000
    p = torch.rand(n, generator=g).clamp(1e-4, 1 - 1e-4)
    p[:8] = torch.tensor([1e-7, 1 - 1e-7, 0.5, 1e-3, 0.999, 0.25, 0.75, 0.9])
    p.requires_grad_(True)
    y = torch.rand(n, generator=g)
    is_pos = (torch.rand(n, generator=g) < 0.3)
    # gfocal_loss.py:33-47 with use_sigmoid=False, element-wise
    loss_neg = F.binary_cross_entropy(p, torch.zeros_like(p), reduction='none') * p.pow(beta)
    loss_pos = F.binary_cross_entropy(p, y, reduction='none') * (y - p).abs().pow(beta)
    ref = torch.where(is_pos, loss_pos, loss_neg)
    ref.sum().backward()
    loss, grad = torch.empty(n), torch.empty(n)
    ip = is_pos.to(torch.int32)
    shim.host_qflp(_p(p.detach()), _p(y), _p(ip), n, C.c_float(beta), _p(loss), _p(grad))
    assert torch.allclose(loss, ref.detach(), rtol=1e-5, atol=1e-7)
    assert float((grad - p.grad).abs().max()) <= 1e-4 * float(p.grad.abs().max())

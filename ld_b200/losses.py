"""Loss modules of the LD path, registered under the reference's names with the reference's
constructor kwargs (mmdet/models/losses/{kd_loss,gfocal_loss,iou_loss}.py).  Inside `LDHead` they are
hyper-parameter carriers for the fused sm_100a loss kernel (`ldb200_ld_loss_fused`); the distillation
losses also run stand-alone through the dense KL kernel.  There is no PyTorch fallback."""
import torch
from torch import nn

from . import ops
from .registry import LOSSES


class _KLDenseFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, pred, soft_label, weight, T, loss_weight, avg_factor):
        loss, d = ops.kl_dense_fwd_bwd(pred.detach().contiguous(), soft_label.detach().contiguous(),
                                       weight.contiguous(), T=T, loss_weight=loss_weight, avg_factor=avg_factor)
        ctx.save_for_backward(d)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        return d * g, None, None, None, None, None


class _MSEFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, target, loss_weight):
        from . import lib as L
        loss = torch.zeros(1, dtype=torch.float32, device=x.device)
        dx = torch.empty_like(x)
        xd = x.detach()
        L.check(L.load().ldb200_mse_fwd_bwd(L.ptr(xd), L.ptr(target), x.numel(), loss_weight, ops._dt(x.dtype),
                                            L.ptr(loss), L.ptr(dx), L.stream_ptr()), 'ldb200_mse_fwd_bwd')
        ctx.save_for_backward(dx)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (dx,) = ctx.saved_tensors
        return dx * g.to(dx.dtype), None, None


@LOSSES.register_module()
class KnowledgeDistillationKLDivLoss(nn.Module):
    """kd_loss.py:39-88: loss_weight * sum_r w_r * mean_bins[KL(softmax(t/T) || softmax(s/T))] * T^2 / avg."""

    def __init__(self, reduction='mean', loss_weight=1.0, T=10):
        super().__init__()
        assert T >= 1
        self.reduction, self.loss_weight, self.T = reduction, loss_weight, T

    def forward(self, pred, soft_label, weight=None, avg_factor=None, reduction_override=None):
        assert reduction_override in (None, 'none', 'mean', 'sum')
        reduction = reduction_override or self.reduction
        assert pred.size() == soft_label.size() and pred.dim() == 2
        if reduction == 'none':
            raise NotImplementedError('reduction="none" is not on the LD train path (ld_head.py:235-271 always reduces)')
        rows = pred.shape[0]
        if weight is None:
            weight = pred.new_ones(rows)
        if avg_factor is None:
            avg = float(rows) if reduction == 'mean' else 1.0
        else:
            if reduction == 'sum':
                raise ValueError('avg_factor can not be used with reduction="sum"')   # losses/utils.py:52-53
            avg = float(avg_factor)
        return _KLDenseFn.apply(pred.float(), soft_label.float(), weight.float(), float(self.T),
                                float(self.loss_weight), max(avg, 1e-12))


@LOSSES.register_module()
class IMLoss(nn.Module):
    """kd_loss.py:99-120 (mse, mean over all elements; weight/avg_factor are ignored by the reference)."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, x, soft_target, weight=None, avg_factor=None, reduction_override=None):
        """Stand-alone path (ldb200_mse_fwd_bwd); inside LDHead the term is part of the fused loss kernel."""
        assert x.shape == soft_target.shape
        return _MSEFn.apply(x.contiguous(), soft_target.detach().to(x.dtype).contiguous(), float(self.loss_weight))


class _HeadLossCfg(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(f'{type(self).__name__} is evaluated inside the fused LDHead loss kernel '
                                  f'(ldb200_ld_loss_fused); no stand-alone / PyTorch path is provided')


@LOSSES.register_module()
class QualityFocalLoss(_HeadLossCfg):
    """gfocal_loss.py:77-135."""

    def __init__(self, use_sigmoid=True, beta=2.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.use_sigmoid, self.beta, self.reduction, self.loss_weight = use_sigmoid, beta, reduction, loss_weight


@LOSSES.register_module()
class DistributionFocalLoss(_HeadLossCfg):
    """gfocal_loss.py:138-179."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight


@LOSSES.register_module()
class GIoULoss(_HeadLossCfg):
    """iou_loss.py:325-360 (eps=1e-6)."""

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight


@LOSSES.register_module()
class CIoULoss(_HeadLossCfg):
    """Named by the teacher config (configs/gfl/gfl_r50_fpn_1x_coco.py:43); the teacher never computes a loss."""

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

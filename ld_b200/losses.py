"""Loss modules of the LD path, registered under the reference's names with the reference's
constructor kwargs (mmdet/models/losses/{kd_loss,gfocal_loss,iou_loss}.py).  Inside `LDHead` they are
hyper-parameter carriers for the fused sm_100a loss kernel (`ldb200_ld_loss_fused`); called on their own
(`build_loss(cfg)(pred, target, weight, avg_factor)`, the registry contract of SURVEY.md 8b) each runs its own
forward+gradient kernel.  There is no PyTorch fallback."""
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


def _reduction_scale(module, rows, weight, avg_factor, reduction_override):
    """losses/utils.py:27-54 `weight_reduce_loss` as (reduction, scale): the kernels compute scale * sum_r w_r * l_r."""
    assert reduction_override in (None, 'none', 'mean', 'sum')
    reduction = reduction_override or module.reduction
    if avg_factor is None:
        scale = 1.0 / max(rows, 1) if reduction == 'mean' else 1.0
    else:
        if reduction == 'sum':
            raise ValueError('avg_factor can not be used with reduction="sum"')      # losses/utils.py:52-53
        scale = 1.0 / float(avg_factor) if reduction == 'mean' else 1.0
    return reduction, float(module.loss_weight) * scale


class _RowLossFn(torch.autograd.Function):
    """A row-wise loss kernel (forward value + d/d pred in one launch).  `launch(row_loss, loss, d_pred)` enqueues it."""

    @staticmethod
    def forward(ctx, pred, launch, reduction):
        rows = pred.shape[0]
        d_pred = torch.empty_like(pred)
        if reduction == 'none':
            out = torch.empty(rows, dtype=torch.float32, device=pred.device)
            launch(out, None, d_pred)
        else:
            acc = torch.zeros(1, dtype=torch.float32, device=pred.device)
            launch(None, acc, d_pred)
            out = acc[0]
        ctx.save_for_backward(d_pred)
        ctx.none = reduction == 'none'
        return out

    @staticmethod
    def backward(ctx, g):
        (d,) = ctx.saved_tensors
        if ctx.none:
            g = g.reshape(-1, *([1] * (d.dim() - 1)))
        return d * g, None, None


def _w(weight, ref):
    return None if weight is None else weight.detach().float().contiguous()


@LOSSES.register_module()
class QualityFocalLoss(nn.Module):
    """gfocal_loss.py:77-135.  forward(pred [N, C], (label [N], score [N]), weight, avg_factor, reduction_override) on
    `ldb200_qfl_fwd_bwd`; `use_sigmoid=False` (every GFLv2 / LDv2 config) takes probabilities (F.binary_cross_entropy
    on joint scores).  Inside LDHead the term is part of the fused loss kernel."""

    def __init__(self, use_sigmoid=True, beta=2.0, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.use_sigmoid, self.beta, self.reduction, self.loss_weight = use_sigmoid, beta, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        from . import lib as L
        label, score = target
        assert pred.dim() == 2 and label.shape == score.shape == pred.shape[:1]
        reduction, scale = _reduction_scale(self, pred.shape[0], weight, avg_factor, reduction_override)
        p = pred.float().contiguous()
        lab, sc, w = label.detach().to(torch.int32).contiguous(), score.detach().float().contiguous(), _w(weight, p)

        def launch(row_loss, loss, d_pred):
            L.check(L.load().ldb200_qfl_fwd_bwd(L.ptr(p.detach()), L.ptr(lab), L.ptr(sc), L.ptr(w), p.shape[0], p.shape[1],
                                                float(self.beta), int(self.use_sigmoid), scale, L.ptr(row_loss),
                                                L.ptr(loss), L.ptr(d_pred), L.stream_ptr()), 'ldb200_qfl_fwd_bwd')
        return _RowLossFn.apply(p, launch, reduction)


@LOSSES.register_module()
class DistributionFocalLoss(nn.Module):
    """gfocal_loss.py:138-179.  forward(pred [N, n+1] logits, target [N] float distances, weight, avg_factor, ...)."""

    def __init__(self, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.reduction, self.loss_weight = reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None):
        from . import lib as L
        assert pred.dim() == 2 and target.shape == pred.shape[:1]
        reduction, scale = _reduction_scale(self, pred.shape[0], weight, avg_factor, reduction_override)
        p, t, w = pred.float().contiguous(), target.detach().float().contiguous(), _w(weight, pred)

        def launch(row_loss, loss, d_pred):
            L.check(L.load().ldb200_dfl_fwd_bwd(L.ptr(p.detach()), L.ptr(t), L.ptr(w), p.shape[0], p.shape[1], scale,
                                                L.ptr(row_loss), L.ptr(loss), L.ptr(d_pred), L.stream_ptr()),
                    'ldb200_dfl_fwd_bwd')
        return _RowLossFn.apply(p, launch, reduction)


@LOSSES.register_module()
class GIoULoss(nn.Module):
    """iou_loss.py:325-360 (eps=1e-6).  forward(pred [N, 4], target [N, 4], weight [N], avg_factor, ...)."""

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

    def forward(self, pred, target, weight=None, avg_factor=None, reduction_override=None, **kwargs):
        from . import lib as L
        assert pred.shape == target.shape and pred.dim() == 2 and pred.shape[1] == 4
        if weight is not None and weight.dim() > 1:
            # iou_loss.py:347-351: a [N, 4] weight (from AnchorHead targets) is reduced to one value per box
            assert weight.shape == pred.shape
            weight = weight.mean(-1)
        reduction, scale = _reduction_scale(self, pred.shape[0], weight, avg_factor, reduction_override)
        p, t, w = pred.float().contiguous(), target.detach().float().contiguous(), _w(weight, pred)

        def launch(row_loss, loss, d_pred):
            L.check(L.load().ldb200_giou_fwd_bwd(L.ptr(p.detach()), L.ptr(t), L.ptr(w), p.shape[0], float(self.eps), scale,
                                                 L.ptr(row_loss), L.ptr(loss), L.ptr(d_pred), L.stream_ptr()),
                    'ldb200_giou_fwd_bwd')
        return _RowLossFn.apply(p, launch, reduction)


class _HeadLossCfg(nn.Module):
    def forward(self, *a, **k):
        raise NotImplementedError(f'{type(self).__name__} is only named by the frozen teacher\'s config; the teacher '
                                  f'never computes a loss (kd_one_stage.py:70-72)')


@LOSSES.register_module()
class CIoULoss(_HeadLossCfg):
    """Named by the teacher config (configs/gfl/gfl_r50_fpn_1x_coco.py:43); the teacher never computes a loss."""

    def __init__(self, eps=1e-6, reduction='mean', loss_weight=1.0):
        super().__init__()
        self.eps, self.reduction, self.loss_weight = eps, reduction, loss_weight

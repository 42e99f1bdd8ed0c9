"""Autograd glue between the reference-shaped nn.Modules and the sm_100a kernels.

Activations are NHWC bf16 tensors `[N, H, W, C]`.  Conv weights are ordinary `nn.Parameter`s of
shape `[Cout, Cin, kh, kw]` (checkpoint-compatible with the reference) kept in channels_last
memory, i.e. physically `[Cout, kh, kw, Cin]` -- which IS the K-major operand layout of the tcgen05
kernels, so packing is a dtype cast and the fp32 weight gradient produced by the wgrad kernel is a
zero-copy view with the parameter's own strides.

Backward of `y = act(scale * conv(x, W) + shift + residual)`:
    g        = dy * (y > 0)                       (relu_bwd kernel)
    d_shift  = colsum(g)                          (conv bias / BN beta)
    G        = wgrad(g, x)                        (tcgen05, fp32)
    dW       = scale[co] * G ,  d_scale[co] = <W[co], G[co]>      (never needs the pre-BN conv output)
    dx       = dgrad(g, (scale*W)^T)              (same tcgen05 kernel, mirrored taps; stride-2 convs as
                                                   4 output-parity classes)
    d_resid  = g  (or its nearest-upsample adjoint when the residual is a coarser FPN map)
"""
import torch

from . import ops


# ------------------------------------------------------------------------------------------------------
# Precision modes.  'bf16' (default, the throughput mode): NHWC bf16 activations, bf16 x bf16 -> fp32 tcgen05 MMAs.
# 'fp32' (the parity mode, SURVEY.md 7.3-1 / north_star "loss and grads within 1e-4 fp32"): NHWC fp32 activations and
# gradients; every convolution / weight gradient runs on split-bf16 operands (x = p0 + p1 + p2, all plane products
# down to 2^-24 accumulated in the fp32 TMEM accumulator), i.e. fp32-accurate results from the same tcgen05 kernels.
_PRECISION = {'mode': 'bf16'}


def set_precision(mode):
    assert mode in ('bf16', 'fp32')
    prev = _PRECISION['mode']
    _PRECISION['mode'] = mode
    return prev


def parity():
    return _PRECISION['mode'] == 'fp32'


def act_dtype():
    return torch.float32 if parity() else torch.bfloat16


class precision:
    """`with engine.precision('fp32'):` -- run the enclosed forward/backward in the fp32-accurate mode."""

    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = set_precision(self.mode)

    def __exit__(self, *a):
        set_precision(self.prev)


def clear_step_cache(params=()):
    """Drop the bf16 operand copies cached on the parameters.  They are keyed by `_version`, which
    torch bumps for in-place torch ops -- but the fused SGD kernel updates the flat parameter buffer
    through raw pointers, so `LDTrainStep` invalidates them explicitly after every step."""
    for p in params:
        for k in ('_ldb_packed', '_ldb_dgrad', '_ldb_packed32', '_ldb_dgrad32'):
            p.__dict__.pop(k, None)


def dense_from_grouped(w, groups):
    """[Cout, Cin/groups, kh, kw] -> block-diagonal [Cout, Cin, kh, kw]: a grouped conv (ResNeXt, resnext.py:64-75) as
    dense tensor-core tiles."""
    co, cig, kh, kw = w.shape
    cog = co // groups
    d = w.new_zeros(co, cig * groups, kh, kw)
    for g in range(groups):
        d[g * cog:(g + 1) * cog, g * cig:(g + 1) * cig] = w[g * cog:(g + 1) * cog]
    return d


def packed_weight(w, groups=1):
    """[Cout,Cin,kh,kw] channels_last fp32 -> [Cout, kh*kw*Cin] bf16 (a cast, no transpose).
    Cached on the parameter (never in an id()-keyed dict: ids are reused after garbage collection)."""
    key = '_ldb_packed32' if parity() else '_ldb_packed'
    hit = getattr(w, key, None)
    if hit is None or hit[0] != w._version or hit[1].device != w.device:
        wd = w.detach() if groups == 1 else dense_from_grouped(w.detach(), groups)
        co, ci, kh, kw = wd.shape
        w2 = wd.permute(0, 2, 3, 1).reshape(co, kh * kw * ci)
        hit = (w._version, ops.split_weight(w2) if parity() else w2.to(torch.bfloat16))
        setattr(w, key, hit)
    return hit[1]


def as_channels_last_param(conv_weight):
    """In-place: give an OIHW parameter channels_last storage (values unchanged)."""
    conv_weight.data = conv_weight.data.contiguous(memory_format=torch.channels_last)
    return conv_weight


def _round64(c):
    return (c + 63) // 64 * 64


def dgrad_weight(w, scale):
    """(scale*W) transposed for the data-gradient conv: [Cin, kh*kw*pad64(Cout)] bf16, zero padded."""
    key = '_ldb_dgrad32' if parity() else '_ldb_dgrad'
    if scale is None:   # level-shared head convs: reuse across the five levels of one backward pass
        hit = getattr(w, key, None)
        if hit is not None and hit[0] == w._version and hit[1][0].device == w.device:
            return hit[1]
    co, ci, kh, kw = w.shape
    ws = w.detach()
    if scale is not None:
        ws = ws * scale.detach().view(-1, 1, 1, 1)
    t = ws.permute(1, 2, 3, 0)                       # [Cin, kh, kw, Cout]
    cop = _round64(co)
    if cop != co:
        t = torch.nn.functional.pad(t, (0, cop - co))
    t = t.reshape(ci, kh * kw * cop)
    out = (ops.split_weight(t) if parity() else t.to(torch.bfloat16).contiguous(), cop)
    if scale is None:
        setattr(w, key, (w._version, out))
    return out


def conv_dgrad(g, w, scale, stride, pad, in_hw, out=None, accumulate=False, mask=None, add=None):
    """dx [N,H,W,Cin] bf16 = conv_transpose(g [N,HO,WO,Cout]) through ldb200_conv2d_tc."""
    co, ci, kh, kw = w.shape
    wd, cop = dgrad_weight(w, scale)
    h, w_ = in_hw
    n = g.shape[0]
    if stride == 1:
        taps = [(pad - s, pad - r, (r * kw + s) * cop) for r in range(kh) for s in range(kw)]
        return ops.conv2d_tc(g, wd, ci, kh, kw, 1, pad, out=out, taps=taps, out_hw=(h, w_),
                             accumulate=accumulate, mask=mask, add=add)
    assert stride == 2 and add is None
    if out is None:   # parity classes write disjoint pixels; untouched ones (1x1 s2) must read as zero
        out = torch.zeros(n, h, w_, ci, dtype=g.dtype, device=g.device)
        accumulate = False
    for py in range(2):
        for px in range(2):
            taps = []
            for r in range(kh):
                if (py + pad - r) % 2:
                    continue
                for s in range(kw):
                    if (px + pad - s) % 2:
                        continue
                    taps.append(((px + pad - s) // 2, (py + pad - r) // 2, (r * kw + s) * cop))
            hp, wp = (h - py + 1) // 2, (w_ - px + 1) // 2
            if not taps or hp <= 0 or wp <= 0:
                continue
            ops.conv2d_tc(g, wd, ci, kh, kw, 1, 0, out=out, taps=taps, out_hw=(hp, wp), flat=False,
                          out_map=((h, w_), (2, py, 2, px)), accumulate=accumulate, mask=mask)
    return out


class OutSlot:
    """Pre-allocated output view (e.g. a level slice of a row-space buffer) handed to a Function as a
    non-tensor argument: autograd then sees an ordinary fresh output, not an in-place write into a view."""

    def __init__(self, t):
        self.t = t


class ConvAffineAct(torch.autograd.Function):
    """y = relu?(scale * conv(x, W) + shift + resample(residual))  -- one tcgen05 launch forward."""

    @staticmethod
    def forward(ctx, x, weight, scale, shift, residual, stride, pad, relu, out, out_dtype, x_premask=False):
        co, ci, kh, kw = weight.shape
        wp = packed_weight(weight)
        y = ops.conv2d_tc(x, wp, co, kh, kw, stride, pad, out=None if out is None else out.t, out_dtype=out_dtype,
                          scale=None if scale is None else scale.detach().contiguous(),
                          bias=None if shift is None else shift.detach().contiguous(),
                          add=residual, relu=relu)
        ctx.cfg = (stride, pad, relu, residual is not None and tuple(residual.shape[1:3]) != tuple(y.shape[1:3]),
                   x_premask)
        ctx.save_for_backward(x, weight, scale, shift, residual, y if relu else None)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, scale, shift, residual, y = ctx.saved_tensors
        stride, pad, relu, resampled, x_premask = ctx.cfg
        co, ci, kh, kw = weight.shape
        need_x, need_w, need_scale, need_shift, need_res = ctx.needs_input_grad[:5]
        if dy.dtype != x.dtype:
            # fp32 head outputs of the bf16 mode: bring the incoming gradient to the activation dtype
            dy = dy.to(x.dtype)
        dy = dy.contiguous()
        g = ops.relu_bwd(dy, y) if relu else dy
        d_x = d_w = d_scale = d_shift = d_res = None
        if need_shift:
            sink = _grad_sink(shift)          # conv bias parameter: accumulate straight into its gradient
            d_shift = sink if sink is not None else zeros_f32(co, g.device)
            ops.colsum(g.reshape(-1, g.shape[-1]), d_shift)
            if sink is not None:
                d_shift = None
        if need_w or need_scale:
            # the wgrad kernel accumulates scale[co]*G straight into the parameter's gradient buffer (the flat
            # gradient buffer of LDTrainStep, or a lazily created channels-last .grad) and emits d_scale from
            # its epilogue: no temporaries, no autograd accumulation kernels
            if weight.grad is None:
                weight.grad = torch.zeros_like(weight, memory_format=torch.channels_last)
            gbuf = weight.grad.permute(0, 2, 3, 1)
            assert gbuf.is_contiguous(), 'conv weight gradients must be channels-last'
            if need_scale:
                d_scale = zeros_f32(co, g.device)
            wm = weight.detach().permute(0, 2, 3, 1).reshape(co, -1)
            ops.conv2d_wgrad_tc(g, x, gbuf.reshape(co, -1), kh, kw, stride, pad,
                                scale=None if scale is None else scale.detach().contiguous(),
                                w_master=wm if need_scale else None, d_scale=d_scale)
        if need_x:
            # x_premask: x is a ResBlockFn output -> hand it a gradient already multiplied by (x > 0)
            d_x = conv_dgrad(g, weight, scale, stride, pad, x.shape[1:3], mask=x if x_premask else None)
        if need_res:
            if resampled:
                d_res = ops.upsample_add_bwd(g, torch.zeros_like(residual))
            else:
                d_res = g
        return d_x, d_w, d_scale, d_shift, d_res, None, None, None, None, None, None


def conv_affine_act(x, weight, scale=None, shift=None, residual=None, stride=1, pad=0, relu=False,
                    out=None, out_dtype=torch.bfloat16, x_premask=False):
    return ConvAffineAct.apply(x, weight, scale, shift, residual, stride, pad, relu,
                               None if out is None else OutSlot(out), out_dtype, x_premask)


class _ZeroPool:
    """Bump allocator over ONE zero-filled fp32 buffer for the small per-channel reduction outputs of the backward
    pass (d_scale, d_shift, column sums): `LDTrainStep` re-zeroes the used extent with one memset per step instead
    of ~270 fill launches.  Inactive outside `zero_pool_begin/.._end` (plain autograd use): `zeros_f32` then falls
    back to `torch.zeros`."""
    buf = None
    off = 0
    active = False


def zero_pool_begin(device, capacity=1 << 20):
    zp = _ZeroPool
    if zp.buf is None or zp.buf.device != device:
        zp.buf, zp.off = torch.zeros(capacity, dtype=torch.float32, device=device), 0
    if zp.off:
        zp.buf[:zp.off].zero_()
    zp.off, zp.active = 0, True


def zero_pool_end():
    _ZeroPool.active = False


def zeros_f32(n, device):
    zp = _ZeroPool
    if zp.active and zp.buf.device == device:
        m = (n + 31) & ~31                      # 128-byte granules keep every view 16-byte aligned
        if zp.off + m <= zp.buf.numel():
            v = zp.buf[zp.off:zp.off + n]
            zp.off += m
            return v
    return torch.zeros(n, dtype=torch.float32, device=device)


def _grad_sink(p):
    """The fp32 `.grad` of a leaf parameter that kernels may accumulate into directly (the flat gradient buffer of
    `LDTrainStep`): saves the zero-fill of a temporary and autograd's AccumulateGrad add.  None -> return the
    gradient through autograd as usual."""
    if p is None or not p.is_leaf or not p.requires_grad:
        return None
    g = p.grad
    if g is None or g.dtype != torch.float32 or not g.is_contiguous() or g.device != p.device:
        return None
    return g


class GNReLU(torch.autograd.Function):
    """y = relu(GroupNorm(x)) on NHWC bf16 (gfl_head.py:102-133 ConvModule norm+act)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps):
        y, stats = ops.gn_relu_fwd(x, gamma.detach().contiguous(), beta.detach().contiguous(), groups, eps, True)
        ctx.save_for_backward(x, y, stats, gamma, beta)
        ctx.cfg = (groups, eps)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, stats, gamma, beta = ctx.saved_tensors
        groups, eps = ctx.cfg
        sg, sb = _grad_sink(gamma), _grad_sink(beta)
        direct = sg is not None and sb is not None
        dg = sg if direct else zeros_f32(gamma.numel(), x.device)
        db = sb if direct else zeros_f32(gamma.numel(), x.device)
        dx = ops.gn_relu_bwd(dy.contiguous(), y, x, stats, gamma.detach().contiguous(), dg, db, groups, eps)
        return (dx, None, None, None, None) if direct else (dx, dg, db, None, None)


def gn_relu(x, gamma, beta, groups=32, eps=1e-5):
    return GNReLU.apply(x, gamma, beta, groups, eps)


class _BNFold(torch.autograd.Function):
    """(gamma, beta) -> (scale, shift) of an eval-mode BatchNorm: one launch forward, one backward."""

    @staticmethod
    def forward(ctx, gamma, beta, mean, var, eps):
        from . import lib as L
        scale, shift = torch.empty_like(gamma), torch.empty_like(gamma)
        L.check(L.load().ldb200_bn_fold(L.ptr(gamma.detach()), L.ptr(beta.detach()), L.ptr(mean), L.ptr(var), float(eps),
                                        L.ptr(scale), L.ptr(shift), gamma.numel(), L.stream_ptr()), 'ldb200_bn_fold')
        ctx.save_for_backward(mean, var, gamma, beta)
        ctx.eps = eps
        return scale, shift

    @staticmethod
    def backward(ctx, d_scale, d_shift):
        from . import lib as L
        mean, var, gamma, beta = ctx.saved_tensors
        for p_ in (gamma, beta):     # accumulate straight into the (flat-buffer) parameter gradients
            if p_.grad is None:
                p_.grad = torch.zeros_like(p_)
        L.check(L.load().ldb200_bn_fold_bwd(L.ptr(None if d_scale is None else d_scale.contiguous()),
                                            L.ptr(None if d_shift is None else d_shift.contiguous()), L.ptr(mean),
                                            L.ptr(var), float(ctx.eps), L.ptr(gamma.grad), L.ptr(beta.grad),
                                            mean.numel(), L.stream_ptr()), 'ldb200_bn_fold_bwd')
        return None, None, None, None, None


def bn_affine(bn):
    """Eval-mode BatchNorm (norm_eval=True, resnet.py:639-649) as a per-channel affine (scale, shift);
    differentiable w.r.t. the (trainable) gamma / beta."""
    frozen = not (bn.weight.requires_grad or bn.bias.requires_grad)
    if frozen:
        key = (bn.weight._version, bn.bias._version, bn.running_mean._version, bn.running_var._version,
               bn.weight.device)
        hit = bn.__dict__.get('_ldb_affine')
        if hit is not None and hit[0] == key:
            return hit[1]
    if not frozen and bn.weight.is_cuda:
        return _BNFold.apply(bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.eps)
    rstd = torch.rsqrt(bn.running_var + bn.eps)
    scale = bn.weight * rstd
    shift = bn.bias - bn.running_mean * scale
    if frozen:
        scale, shift = scale.detach().contiguous(), shift.detach().contiguous()
        bn.__dict__['_ldb_affine'] = (key, (scale, shift))
    return scale, shift


# ------------------------------------------------------------------------------------------------------
# Residual blocks as ONE autograd node each, so that the backward chain can use the fused dgrad epilogues:
# ReLU masks are applied by the dgrad that PRODUCES a gradient (mask = the forward activation it flows into) and
# the identity / downsample branch is added in the same epilogue -- no relu_bwd and no add kernels.
#
# Convention between blocks ("pre-masked gradients"): the gradient a block RECEIVES for its output has already
# been multiplied by (out > 0), and the gradient it RETURNS for its input has been multiplied by (x > 0).  Every
# consumer of a block output must follow it (next block, FPN lateral conv with `x_premask=True`).
def _wgrad_into(weight, g, x, scale, stride, pad, need_scale):
    co, ci, kh, kw = weight.shape
    if weight.grad is None:
        weight.grad = torch.zeros_like(weight, memory_format=torch.channels_last)
    gbuf = weight.grad.permute(0, 2, 3, 1)
    assert gbuf.is_contiguous(), 'conv weight gradients must be channels-last'
    d_scale = zeros_f32(co, g.device) if need_scale else None
    ops.conv2d_wgrad_tc(g, x, gbuf.reshape(co, -1), kh, kw, stride, pad,
                        scale=None if scale is None else scale.detach().contiguous(),
                        w_master=weight.detach().permute(0, 2, 3, 1).reshape(co, -1) if need_scale else None,
                        d_scale=d_scale)
    return d_scale


def _colsum(g):
    out = zeros_f32(g.shape[-1], g.device)
    ops.colsum(g.reshape(-1, g.shape[-1]), out)
    return out


class ResBlockFn(torch.autograd.Function):
    """Bottleneck (3 convs) or BasicBlock (2 convs) with eval-BN folded into per-channel scale/shift.

    args: x, then for every conv (weight, scale, shift) in order [main convs..., downsample?]."""

    @staticmethod
    def forward(ctx, x, strides, pads, has_ds, *params):
        nmain = len(strides)
        convs = [params[3 * i:3 * i + 3] for i in range(nmain + (1 if has_ds else 0))]
        stride = max(strides)
        acts = [x]
        idt = x
        if has_ds:
            w, s, b = convs[nmain]
            idt = ops.conv2d_tc(x, packed_weight(w), w.shape[0], 1, 1, stride, 0, scale=s.detach().contiguous(),
                                bias=b.detach().contiguous())
        h = x
        for i in range(nmain):
            w, s, b = convs[i]
            last = i == nmain - 1
            h = ops.conv2d_tc(h, packed_weight(w), w.shape[0], w.shape[2], w.shape[3], strides[i], pads[i],
                              scale=s.detach().contiguous(), bias=b.detach().contiguous(),
                              add=idt if last else None, relu=True)
            acts.append(h)
        ctx.cfg = (strides, pads, has_ds)
        ctx.save_for_backward(*acts[:nmain], *params)   # the block output itself is not needed (pre-masked grads)
        return h

    @staticmethod
    def backward(ctx, g):
        strides, pads, has_ds = ctx.cfg
        nmain = len(strides)
        saved = ctx.saved_tensors
        acts, params = saved[:nmain], saved[nmain:]
        convs = [params[3 * i:3 * i + 3] for i in range(nmain + (1 if has_ds else 0))]
        x = acts[0]
        need_x = ctx.needs_input_grad[0]
        grads = [None] * len(params)
        g = g.contiguous()          # grad w.r.t. the block output, pre-masked by (out > 0)
        g_out = g
        for i in range(nmain - 1, -1, -1):
            w, s, b = convs[i]
            a_in = acts[i]
            ns, nb = ctx.needs_input_grad[4 + 3 * i + 1], ctx.needs_input_grad[4 + 3 * i + 2]
            if nb:
                grads[3 * i + 2] = _colsum(g)
            if ctx.needs_input_grad[4 + 3 * i] or ns:
                grads[3 * i + 1] = _wgrad_into(w, g, a_in, s, strides[i], pads[i], ns)
            if i > 0:
                g = conv_dgrad(g, w, s, strides[i], pads[i], a_in.shape[1:3], mask=a_in)
            elif need_x:
                if has_ds:
                    wd, sd, bd = convs[nmain]
                    t = conv_dgrad(g_out, wd, sd, max(strides), 0, x.shape[1:3])
                    if strides[0] == 1:
                        g = conv_dgrad(g, w, s, 1, pads[0], x.shape[1:3], add=t, mask=x)
                    else:   # BasicBlock: the strided 3x3 covers every pixel through its four parity classes
                        g = conv_dgrad(g, w, s, strides[0], pads[0], x.shape[1:3], out=t, accumulate=True, mask=x)
                else:
                    g = conv_dgrad(g, w, s, strides[0], pads[0], x.shape[1:3], add=g_out, mask=x)
            else:
                g = None
        if has_ds:
            wd, sd, bd = convs[nmain]
            j = nmain
            ns, nb = ctx.needs_input_grad[4 + 3 * j + 1], ctx.needs_input_grad[4 + 3 * j + 2]
            if nb:
                grads[3 * j + 2] = _colsum(g_out)
            if ctx.needs_input_grad[4 + 3 * j] or ns:
                grads[3 * j + 1] = _wgrad_into(wd, g_out, x, sd, max(strides), 0, ns)
        return (g, None, None, None, *grads)


def res_block(x, main, downsample=None):
    """main: list of (conv, bn, stride, pad); downsample: (conv, bn) or None."""
    params, strides, pads = [], [], []
    for conv, bn, st, pd in main:
        s, b = bn_affine(bn)
        params += [conv._cl(), s, b]
        strides.append(st)
        pads.append(pd)
    if downsample is not None:
        s, b = bn_affine(downsample[1])
        params += [downsample[0]._cl(), s, b]
    return ResBlockFn.apply(x, tuple(strides), tuple(pads), downsample is not None, *params)


class ConvLevelsFn(torch.autograd.Function):
    """A level-shared 3x3 head conv (no bias / norm handled by the caller) on all FPN levels in ONE tcgen05 launch
    forward and ONE for the data gradients; weight gradients accumulate level by level into `weight.grad`."""

    @staticmethod
    def forward(ctx, weight, pad, *xs):
        co, ci, kh, kw = weight.shape
        ys = ops.conv2d_tc_multi(list(xs), packed_weight(weight), co, kh, kw, 1, pad)
        ctx.pad = pad
        ctx.save_for_backward(weight, *xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gs):
        weight, xs = ctx.saved_tensors[0], ctx.saved_tensors[1:]
        pad = ctx.pad
        co, ci, kh, kw = weight.shape
        gs = [g.contiguous() for g in gs]
        if ctx.needs_input_grad[0]:
            for g, x in zip(gs, xs):
                _wgrad_into(weight, g, x, None, 1, pad, False)
        dxs = [None] * len(xs)
        if any(ctx.needs_input_grad[2:]):
            wd, cop = dgrad_weight(weight, None)
            taps = [(pad - s, pad - r, (r * kw + s) * cop) for r in range(kh) for s in range(kw)]
            dxs = ops.conv2d_tc_multi(gs, wd, ci, kh, kw, 1, pad, taps=taps)
        return (None, None, *dxs)


def conv_levels(weight, xs, pad=1):
    return list(ConvLevelsFn.apply(weight, pad, *xs))


import os as _os
_GN_READ_Y = _os.environ.get('LDB200_GN_READ_Y') is not None


class ConvGNReLULevelsFn(torch.autograd.Function):
    """One GFL tower layer (gfl_head.py:102-133, 177-180: level-shared 3x3 conv without bias -> GroupNorm(32) -> ReLU)
    on ALL pyramid levels as one autograd node:
      forward : 1 tcgen05 launch for the five convs, its epilogue also accumulates the GroupNorm statistics
                (no pass over the conv output for them), 1 normalise+ReLU launch for the five levels;
      backward: 2 GroupNorm launches (reductions, apply) for the five levels, 5 weight-gradient launches, 1 data-gradient
                launch.  (Before: 5 x (stats, apply) forward and 5 x (reductions, apply) backward launches per layer.)"""

    @staticmethod
    def forward(ctx, weight, gamma, beta, pad, groups, eps, *xs):
        co, ci, kh, kw = weight.shape
        dev = xs[0].device
        sizes = [x.shape[0] * groups * 2 for x in xs]
        flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)      # one fill for the stats of all levels
        stats, o = [], 0
        for k in sizes:
            stats.append(flat[o:o + k])
            o += k
        cs = ops.conv2d_tc_multi(list(xs), packed_weight(weight), co, kh, kw, 1, pad, gn_stats=stats)
        ys = ops.gn_apply_multi(cs, stats, gamma.detach().contiguous(), beta.detach().contiguous(), groups, eps, True)
        ctx.cfg = (pad, groups, eps, len(xs))
        ctx.save_for_backward(weight, gamma, beta, flat, *xs, *cs, *ys)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        pad, groups, eps, nl = ctx.cfg
        saved = ctx.saved_tensors
        weight, gamma, beta, flat = saved[:4]
        xs, cs, ys = saved[4:4 + nl], saved[4 + nl:4 + 2 * nl], saved[4 + 2 * nl:4 + 3 * nl]
        co, ci, kh, kw = weight.shape
        stats, o = [], 0
        for x in xs:
            k = x.shape[0] * groups * 2
            stats.append(flat[o:o + k])
            o += k
        sg, sb = _grad_sink(gamma), _grad_sink(beta)
        direct = sg is not None and sb is not None
        dg = sg if direct else zeros_f32(gamma.numel(), xs[0].device)
        db = sb if direct else zeros_f32(gamma.numel(), xs[0].device)
        dys = [d.contiguous() if d.dtype == cs[0].dtype else d.to(cs[0].dtype).contiguous() for d in dys]
        # the ReLU mask is recomputed from the conv output and the statistics: y is not read (LDB200_GN_READ_Y=1: read it)
        gs = ops.gn_relu_bwd_multi(dys, ys if _GN_READ_Y else None, cs, stats, gamma.detach().contiguous(), dg, db,
                                   groups, eps, beta=beta.detach().contiguous())
        if ctx.needs_input_grad[0]:
            for g, x in zip(gs, xs):
                _wgrad_into(weight, g, x, None, 1, pad, False)
        dxs = [None] * nl
        if any(ctx.needs_input_grad[6:]):
            wd, cop = dgrad_weight(weight, None)
            taps = [(pad - s, pad - r, (r * kw + s) * cop) for r in range(kh) for s in range(kw)]
            dxs = ops.conv2d_tc_multi(gs, wd, ci, kh, kw, 1, pad, taps=taps)
        return (None, None if direct else dg, None if direct else db, None, None, None, *dxs)


def conv_gn_relu_levels(weight, gamma, beta, xs, pad=1, groups=32, eps=1e-5):
    return list(ConvGNReLULevelsFn.apply(weight, gamma, beta, pad, groups, eps, *xs))

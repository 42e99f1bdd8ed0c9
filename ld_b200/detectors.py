"""Detectors registered under the reference's names: `GFL` (the teacher's type) and
`KnowledgeDistillationSingleStageDetector` (mmdet/models/detectors/{base,single_stage,gfl,
kd_one_stage}.py), plus the data-parallel train step (flat gradient buffer, ONE NCCL all-reduce, fused
SGD) that replaces mmcv's MMDistributedDataParallel + OptimizerHook + torch.optim.SGD
(mmdet/apis/train.py:74-127)."""
import os
from collections import OrderedDict

import torch
import torch.distributed as dist
from torch import nn

from . import engine, lib as L, ops
from .config import Config
from .registry import DETECTORS, build_backbone, build_head, build_neck


# The reference loads `pretrained` (torchvision://resnetNN, downloaded) and `teacher_ckpt` through mmcv's
# load_checkpoint, which raises on a missing file (kd_one_stage.py:40-44, single_stage.py:35-50).  There is no network
# here, so a checkpoint that cannot be found is an ERROR unless random initialisation is explicitly allowed
# (synthetic-data benchmarks and tests): `with ld_b200.detectors.allow_random_init(): build_detector(cfg)` or
# LDB200_ALLOW_RANDOM_INIT=1.
_RANDOM_INIT = {'ok': False}


def _random_ok():
    return _RANDOM_INIT['ok'] or os.environ.get('LDB200_ALLOW_RANDOM_INIT') is not None


class allow_random_init:
    def __enter__(self):
        self.prev = _RANDOM_INIT['ok']
        _RANDOM_INIT['ok'] = True

    def __exit__(self, *a):
        _RANDOM_INIT['ok'] = self.prev


def _find_checkpoint(spec):
    """A local file for `spec`: a path, or torchvision://NAME looked up under $LD_B200_PRETRAINED_DIR and torch hub's
    checkpoint cache (NAME*.pth).  None when nothing is found."""
    spec = str(spec)
    if os.path.isfile(spec):
        return spec
    if spec.startswith('torchvision://'):
        import glob
        name = spec.split('://', 1)[1]
        roots = [os.environ.get('LD_B200_PRETRAINED_DIR', ''), os.path.join(torch.hub.get_dir(), 'checkpoints')]
        for r in roots:
            hits = sorted(glob.glob(os.path.join(r, name + '*.pth'))) if r else []
            if hits:
                return hits[0]
    return None


def load_checkpoint_into(module, spec, what, prefix_candidates=('',)):
    """mmcv.runner.load_checkpoint semantics: missing file -> error (unless random init is allowed); missing /
    unexpected keys are reported, not hidden."""
    import warnings
    path = _find_checkpoint(spec)
    if path is None:
        if _random_ok():
            warnings.warn(f'{what}: checkpoint {spec!r} not found -- keeping the RANDOM initialisation '
                          f'(allowed by allow_random_init / LDB200_ALLOW_RANDOM_INIT)')
            return False
        raise FileNotFoundError(f'{what}: checkpoint {spec!r} not found (no network on this box: put the file under '
                                f'$LD_B200_PRETRAINED_DIR, or opt into random init with '
                                f'ld_b200.detectors.allow_random_init() / LDB200_ALLOW_RANDOM_INIT=1)')
    ckpt = torch.load(path, map_location='cpu', weights_only=False)
    sd = ckpt.get('state_dict', ckpt) if isinstance(ckpt, dict) else ckpt
    res = module.load_state_dict(sd, strict=False)
    if res.missing_keys or res.unexpected_keys:
        warnings.warn(f'{what}: loaded {path} with {len(res.missing_keys)} missing and {len(res.unexpected_keys)} '
                      f'unexpected keys; missing: {res.missing_keys[:8]} unexpected: {res.unexpected_keys[:8]}')
    return True


class BaseDetector(nn.Module):
    """detectors/base.py: forward dispatch, train_step, _parse_losses."""

    def forward(self, img, img_metas, return_loss=True, **kwargs):
        if return_loss:
            return self.forward_train(img, img_metas, **kwargs)
        return self.forward_test(img, img_metas, **kwargs)

    def forward_test(self, imgs, img_metas, **kwargs):
        """base.py:121-167: lists of images / metas (one entry per test-time augmentation)."""
        if isinstance(imgs, torch.Tensor):
            imgs, img_metas = [imgs], [img_metas]
        if len(imgs) != 1:
            raise NotImplementedError('aug_test (multi-scale / flip test-time augmentation) is outside the LD path')
        return self.simple_test(imgs[0], img_metas[0], **kwargs)

    def _parse_losses(self, losses):
        """base.py:185-218: loss = sum over keys containing 'loss' of the per-level sums; one device->host
        copy for the 9+1 log scalars instead of ten allreduce+.item() pairs."""
        log_vars = OrderedDict()
        for name, value in losses.items():
            if isinstance(value, torch.Tensor):
                log_vars[name] = value.mean()
            elif isinstance(value, list):
                log_vars[name] = sum(v.mean() for v in value)
            else:
                raise TypeError(f'{name} is not a tensor or list of tensors')
        loss = sum(v for k, v in log_vars.items() if 'loss' in k)
        log_vars['loss'] = loss
        keys = list(log_vars)
        vec = torch.stack([log_vars[k].detach().float() for k in keys])
        if dist.is_available() and dist.is_initialized():
            dist.all_reduce(vec)
            vec = vec / dist.get_world_size()
        host = vec.cpu()
        return loss, OrderedDict((k, float(host[i])) for i, k in enumerate(keys))

    def train_step(self, data, optimizer):
        """base.py:220-253."""
        losses = self(**data)
        loss, log_vars = self._parse_losses(losses)
        return dict(loss=loss, log_vars=log_vars, num_samples=len(data['img_metas']))


@DETECTORS.register_module()
class SingleStageDetector(BaseDetector):
    """detectors/single_stage.py:17-57."""

    def __init__(self, backbone, neck=None, bbox_head=None, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__()
        self.backbone = build_backbone(backbone)
        if neck is not None:
            self.neck = build_neck(neck)
        bbox_head = dict(bbox_head)
        bbox_head.update(train_cfg=train_cfg, test_cfg=test_cfg)
        self.bbox_head = build_head(bbox_head)
        self.train_cfg, self.test_cfg = train_cfg, test_cfg
        self.init_weights(pretrained=pretrained)

    @property
    def with_neck(self):
        return hasattr(self, 'neck') and self.neck is not None

    def init_weights(self, pretrained=None):
        """single_stage.py:35-50.  A `pretrained` string is resolved to a local file (no download here); a file that
        cannot be found raises unless random init is allowed."""
        self.backbone.init_weights(pretrained=None)
        if self.with_neck:
            self.neck.init_weights()
        self.bbox_head.init_weights()
        if isinstance(pretrained, str):
            load_checkpoint_into(self.backbone, pretrained, f'{type(self).__name__}.backbone pretrained')

    def simple_test(self, img, img_metas, rescale=False):
        """single_stage.py:96-121: backbone + neck + head forward, `get_bboxes` (top-k, decode, NMS on the GPU),
        `bbox2result` (core/bbox/transforms.py:92-116): per image a list of num_classes numpy arrays [k, 5]."""
        with torch.no_grad():
            x = self.extract_feat(img)
            head = self.bbox_head
            outs = head(x, with_scores=True) if hasattr(head, 'reg_conf') else head(x)   # GFocalHead: joint scores
            bbox_list = head.get_bboxes(outs[0], outs[1], img_metas, rescale=rescale)
        nc = self.bbox_head.num_classes
        results = []
        for det, lab in bbox_list:
            det, lab = det.cpu().numpy(), lab.cpu().numpy()
            results.append([det[lab == c] for c in range(nc)])
        return results

    def extract_feat(self, img, im2col=None, stem_out=None):
        x = self.backbone(img, im2col=im2col) if stem_out is None else self.backbone(img, stem_out=stem_out)
        if self.with_neck:
            x = self.neck(x)
        return x


@DETECTORS.register_module()
class GFL(SingleStageDetector):
    """detectors/gfl.py."""


def _resolve_config(path):
    for root in ('', os.environ.get('LD_B200_CONFIG_ROOT', ''), '/root/reference'):
        p = os.path.join(root, path) if root else path
        if os.path.isfile(p):
            return p
    raise FileNotFoundError(f'teacher_config {path} not found (relative to CWD, $LD_B200_CONFIG_ROOT)')


@DETECTORS.register_module()
class KnowledgeDistillationSingleStageDetector(SingleStageDetector):
    """kd_one_stage.py:12-108."""

    def __init__(self, backbone, neck, bbox_head, teacher_config, output_feature=False, teacher_ckpt=None,
                 eval_teacher=True, train_cfg=None, test_cfg=None, pretrained=None):
        super().__init__(backbone, neck, bbox_head, train_cfg, test_cfg, pretrained)
        self.eval_teacher = eval_teacher
        self.output_feature = output_feature
        if isinstance(teacher_config, str):
            teacher_config = Config.fromfile(_resolve_config(teacher_config))
        from .registry import build_detector
        tmodel = teacher_config['model']
        tmodel = tmodel.to_dict() if hasattr(tmodel, 'to_dict') else dict(tmodel)
        from .config import ConfigDict
        self.teacher_model = build_detector(ConfigDict(tmodel))
        if teacher_ckpt is not None:
            load_checkpoint_into(self.teacher_model, teacher_ckpt, 'teacher_ckpt')
        for p in self.teacher_model.parameters():
            p.requires_grad = False

    def forward_train(self, img, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore=None):
        """kd_one_stage.py:46-81.  The stem's im2col operand is built once and shared by student and teacher."""
        if not self.output_feature:
            raise NotImplementedError('LDHead needs output_feature=True (kd_one_stage.py:73-76 mis-binds its '
                                      'arguments otherwise; both COCO LD configs set it)')
        # both stems (student + teacher) in ONE tcgen05 launch with the im2col done in shared memory: no 1.65 GB im2col
        # buffer, 1.02 ms instead of 1.40 + 2 x 0.35 ms (bench: 46.0 vs 47.4 ms / step).  LDB200_NO_FUSED_STEM=1: the
        # im2col + GEMM route (also what the fp32-accurate mode uses)
        fused_stem = os.environ.get('LDB200_NO_FUSED_STEM') is None and not engine.parity()
        if fused_stem:
            ws, afs = self.backbone.stem_params()
            wt, aft = self.teacher_model.backbone.stem_params()
            stem_s, stem_t = ops.stem_tc(img.float().contiguous(), torch.cat([ws, wt]).contiguous(), [afs, aft])
            return self._forward_train_fused_stem(img, stem_s, stem_t, img_metas, gt_bboxes, gt_labels,
                                                  gt_bboxes_ignore)
        im2col = ops.stem_im2col(img.float().contiguous(), engine.act_dtype())
        # the frozen teacher and the student only share the input: run the teacher on a second stream so its
        # kernels fill the SMs the student's tails / latency-bound epilogues leave idle (also captured as a
        # fork/join inside the CUDA graph)
        main = torch.cuda.current_stream()
        side = self._teacher_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side), torch.no_grad():
            teacher_x = self.teacher_model.extract_feat(img, im2col=im2col)
            out_teacher = self.teacher_model.bbox_head(teacher_x)
        x = self.extract_feat(img, im2col=im2col)
        main.wait_stream(side)
        # (LDHead: (cls_scores, bbox_preds); LDv2Head: (None, bbox_preds, cls_feats))
        for t in list(teacher_x) + [t for grp in out_teacher if grp is not None for t in grp] + [im2col[0]]:
            t.record_stream(main)
        return self.bbox_head.forward_train(x, out_teacher, teacher_x, img_metas, gt_bboxes, gt_labels,
                                            gt_bboxes_ignore)

    def _forward_train_fused_stem(self, img, stem_s, stem_t, img_metas, gt_bboxes, gt_labels, gt_bboxes_ignore):
        main = torch.cuda.current_stream()
        side = self._teacher_stream()
        side.wait_stream(main)
        with torch.cuda.stream(side), torch.no_grad():
            teacher_x = self.teacher_model.extract_feat(img, stem_out=stem_t)
            out_teacher = self.teacher_model.bbox_head(teacher_x)
        x = self.extract_feat(img, stem_out=stem_s)
        main.wait_stream(side)
        for t in list(teacher_x) + [t for grp in out_teacher if grp is not None for t in grp]:
            t.record_stream(main)
        stem_t.record_stream(side)      # allocated on the main stream, read by the teacher's pool on the side stream
        return self.bbox_head.forward_train(x, out_teacher, teacher_x, img_metas, gt_bboxes, gt_labels,
                                            gt_bboxes_ignore)

    def _teacher_stream(self):
        dev = torch.cuda.current_device()
        st = self.__dict__.get('_side_stream')
        if st is None or st.device.index != dev:
            st = torch.cuda.Stream(device=dev)
            self.__dict__['_side_stream'] = st
        return st

    def cuda(self, device=None):
        self.teacher_model.cuda(device=device)
        return super().cuda(device=device)

    def to(self, *args, **kwargs):
        self.teacher_model.to(*args, **kwargs)
        return super().to(*args, **kwargs)

    def train(self, mode=True):
        if self.eval_teacher:
            self.teacher_model.train(False)
        else:
            self.teacher_model.train(mode)
        super().train(mode)
        return self

    def __setattr__(self, name, value):
        """The teacher is a plain attribute: absent from parameters()/state_dict()/the gradient
        all-reduce (kd_one_stage.py:97-108)."""
        if name == 'teacher_model':
            object.__setattr__(self, name, value)
        else:
            super().__setattr__(name, value)


class LDTrainStep:


    """One data-parallel LD training iteration on B200:

        zero flat grad -> forward (student + teacher) -> fused loss -> backward -> ONE NCCL all-reduce of
        the flat fp32 gradient buffer -> fused SGD(momentum, weight decay, linear warm-up) on the flat
        parameter buffer.

    Replaces MMDistributedDataParallel's bucketed all-reduce, mmcv OptimizerHook, LrUpdaterHook
    (warm-up) and torch.optim.SGD (configs/ld/ld_r50_gflv1_r101_fpn_coco_1x.py:53,
    configs/_base_/schedules/schedule_1x.py)."""

    lr_steps, lr_gamma = (), 0.1      # class defaults: no step decay unless the constructor is given `lr_steps`

    def __init__(self, model, lr=0.01, momentum=0.9, weight_decay=1e-4, warmup_iters=500, warmup_ratio=0.001,
                 precision='bf16', lr_steps=(), lr_gamma=0.1):
        """precision: 'bf16' = throughput mode (bf16 activations, single-pass tcgen05 MMAs); 'fp32' = parity mode
        (fp32 activations, split-bf16 convolutions: matches the reference's fp32 arithmetic to ~1e-6 per conv)."""
        assert precision in ('bf16', 'fp32')
        self.precision = precision
        # mmcv StepLrUpdaterHook (configs/_base_/schedules/schedule_1x.py: step=[8, 11] epochs): lr_steps are given in
        # ITERATIONS here (epoch * iterations_per_epoch), lr = base * gamma ** (number of steps passed)
        self.lr_steps, self.lr_gamma = tuple(sorted(lr_steps)), lr_gamma
        self.model = model
        self.base_lr, self.momentum, self.wd = lr, momentum, weight_decay
        self.warmup_iters, self.warmup_ratio = warmup_iters, warmup_ratio
        self.iter = 0
        params = [p for p in model.parameters() if p.requires_grad]
        self.params = params
        n = sum(p.numel() for p in params)
        dev = params[0].device
        self.flat_w = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat_g = torch.zeros(n, dtype=torch.float32, device=dev)
        self.flat_m = torch.zeros(n, dtype=torch.float32, device=dev)
        o = 0
        for p in params:
            k = p.numel()
            if p.dim() == 4:   # conv weight: channels-last storage == packed operand layout
                co, ci, kh, kw = p.shape
                seg = self.flat_w[o:o + k].view(co, kh, kw, ci)
                seg.copy_(p.data.permute(0, 2, 3, 1))
                p.data = seg.permute(0, 3, 1, 2)
                p.grad = self.flat_g[o:o + k].view(co, kh, kw, ci).permute(0, 3, 1, 2)
            else:
                seg = self.flat_w[o:o + k].view(p.shape)
                seg.copy_(p.data)
                p.data = seg
                p.grad = self.flat_g[o:o + k].view(p.shape)
            o += k
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        if self.world > 1:
            dist.broadcast(self.flat_w, src=0)
        self._init_overlap()

    # ------------------------------------------------------------------ overlapped gradient exchange
    # DDP semantics (mmdet/apis/train.py:78-82: bucketed all-reduce overlapped with backward): the flat gradient
    # buffer is exchanged in segments ordered by BACKWARD COMPLETION -- head, neck, layer4, layer3, rest -- each on a
    # communication stream as soon as the last kernel writing into it has been issued, so only the last segment's
    # exchange (the first backbone stage that trains) is exposed after the backward pass.  The parameter order of
    # `model.parameters()` (backbone, neck, bbox_head) makes every segment one contiguous slice of flat_g.
    SEGMENT_PREFIXES = ('bbox_head.', 'neck.', 'backbone.layer4.', 'backbone.layer3.')

    def _init_overlap(self):
        names = [k for k, p in self.model.named_parameters() if p.requires_grad]
        assert len(names) == len(self.params)
        spans, o = {}, 0
        for k, p in zip(names, self.params):
            for pre in self.SEGMENT_PREFIXES:
                if k.startswith(pre):
                    a, b = spans.get(pre, (o, o))
                    assert b == o or a == b, f'parameters of {pre} are not contiguous in the flat buffer'
                    spans[pre] = (a if b > a else o, o + p.numel())
                    break
            o += p.numel()
        self._segments = spans            # prefix -> (start, end) in flat_g
        self._reduced = []                # slices already exchanged in the current iteration
        # opt-in (LDB200_AR_OVERLAP=1): the exchange is ~0.4 ms of a 46 ms step on NVLink 5, and the default path (one
        # all-reduce after the backward pass, captured in the graph) is the one measured at N = 2..8
        self.overlap = os.environ.get('LDB200_AR_OVERLAP') == '1' and os.environ.get('LDB200_NO_AR_OVERLAP') is None
        self._hook_handle = None
        if self.world > 1 and self.overlap and hasattr(self.model, 'neck'):
            self._hook_handle = self.model.neck.register_forward_hook(self._neck_forward_hook)

    def _neck_forward_hook(self, module, inputs, outputs):
        """Remember the stage-boundary tensors of THIS forward pass (backbone outputs, FPN outputs)."""
        if self._in_step and torch.is_grad_enabled() and self._staged_now():
            self._stage = dict(fpn=list(outputs), feats=list(inputs[0]))

    def _backward(self, loss):
        """`loss.backward()`, or -- data-parallel with the overlapped exchange -- the same backward pass cut at the stage
        boundaries (loss -> FPN outputs -> backbone outputs C5 -> C4 -> C3 -> rest): after each piece the segment of
        the flat gradient buffer it completed goes to the communication stream, while the next piece computes.  All
        from the calling thread; the gradients are the ones a single backward call produces (the cuts only order the
        work; tests/test_gpu_staged_backward.py)."""
        st = getattr(self, '_stage', None)
        if not (self._staged_now() and st is not None):
            loss.backward()
            return
        log = getattr(self, '_cut_log', None)      # debugging aid: the gradients at the cuts
        self._stage = None
        named = [(k, p) for k, p in self.model.named_parameters() if p.requires_grad]
        seg_params = {pre: [p for k, p in named if k.startswith(pre)] for pre in self.SEGMENT_PREFIXES}
        rest = [p for k, p in named if not k.startswith(self.SEGMENT_PREFIXES)]
        fpn = [t for t in st['fpn'] if t.requires_grad]
        # Where to cut: `torch.autograd.backward(inputs=[t])` also RUNS the node that produced t, so every cut tensor
        # must come out of a node whose backward does no work: the NCHW permute views the backbone hands the neck
        # (their gradient is the part that arrives through the FPN laterals) and the aliases of the stage outputs the
        # backbone feeds its next stage (`ResNet._cuts`; their gradient is the part that comes down the backbone).
        views = [t for t in st['feats'] if t.requires_grad][::-1]             # C5, C4, C3 as the neck saw them
        cuts = list(getattr(self.model.backbone, '_cuts', ()))[::-1]          # inputs of layer4, layer3, ...
        self.model.backbone._cuts = []
        aligned = bool(views) and all(v._is_view() for v in views) and len(cuts) >= len(views) - 1 and \
            all(views[k + 1].data_ptr() == cuts[k].data_ptr() for k in range(len(views) - 1))
        if not aligned:            # another backbone / out_indices: one piece
            loss.backward()
            return
        # (retain_graph: without it the engine also releases the saved tensors of the nodes BELOW the cut)
        torch.autograd.backward(loss, inputs=fpn + seg_params['bbox_head.'], retain_graph=True)
        self._reduce_span('bbox_head.')
        g_fpn = [t.grad for t in fpn]
        for t in fpn:
            t.grad = None
        torch.autograd.backward(fpn, g_fpn, inputs=views + seg_params['neck.'], retain_graph=True)
        self._reduce_span('neck.')
        lateral = [t.grad for t in views]
        for t in views:
            t.grad = None
        if log is not None:
            log.update({f'lateral{k}': v.clone() for k, v in enumerate(lateral) if v is not None})
        root, g = views[0], lateral[0]
        cur = 0
        for name in ('backbone.layer4.', 'backbone.layer3.'):
            if cur + 1 >= len(views):
                break
            c = cuts[cur]
            torch.autograd.backward([root], [g], inputs=[c] + seg_params[name], retain_graph=True)
            self._reduce_span(name)
            cur += 1
            if log is not None:
                log[f'through{cur}'] = c.grad.clone()
            # gradient w.r.t. the view -> layout of the stage output it permutes
            root, g = c, lateral[cur].permute(0, 2, 3, 1) + c.grad
            c.grad = None
        tail = rest + [p for name in self.SEGMENT_PREFIXES[2 + cur:] for p in seg_params[name]]
        if tail:
            torch.autograd.backward([root], [g], inputs=tail)

    def _staged_now(self):
        """Backward in pieces with the exchange overlapped?  Data parallel, overlap not disabled, and not inside
        `capture()`: replayed from a CUDA graph the segment exchanges on the communication stream gave wrong sums on
        2 GPUs (eager: validated against the single exchange), so the captured iteration keeps the single all-reduce
        after the backward pass unless LDB200_GRAPH_OVERLAP=1."""
        if not ((self.world > 1 or getattr(self, '_force_staged', False)) and self.overlap):
            return False
        return not getattr(self, '_graph_mode', False) or os.environ.get('LDB200_GRAPH_OVERLAP') == '1'

    def _comm(self):
        st = self.__dict__.get('_comm_stream')
        if st is None:
            st = torch.cuda.Stream(device=self.flat_g.device)
            self.__dict__['_comm_stream'] = st
        return st

    def _reduce_span(self, seg):
        """All-reduce one segment of flat_g on the communication stream, after everything issued so far on the compute
        streams (main + the head's regression-tower stream)."""
        if self.world == 1 or seg in self._done or seg not in self._segments:
            return
        self._done.add(seg)
        a, b = self._segments[seg]
        if b <= a:
            return
        comm = self._comm()
        comm.wait_stream(torch.cuda.current_stream())
        side = self.model.bbox_head.__dict__.get('_side_stream')
        if side is not None:
            comm.wait_stream(side)
        with torch.cuda.stream(comm):
            dist.all_reduce(self.flat_g[a:b])
        self._reduced.append((a, b))

    def _begin_exchange(self):
        self._done, self._reduced, self._in_step = set(), [], True
        bb = getattr(self.model, 'backbone', None)
        if bb is not None and hasattr(bb, 'cut_stages'):
            bb.cut_stages = self._staged_now()
        # the step's loss is the plain sum of all terms (base.py:185-218): the loss node may hand out the gradients it
        # precomputed in the forward pass without checking its upstream gradient (a device->host read)
        self.model.bbox_head._unit_loss_grad = True

    def _finish_exchange(self):
        """Exchange whatever the hooks have not covered (the earliest training stage; everything when overlap is off),
        then make the main stream wait for the communication stream."""
        self._in_step = False
        self.model.bbox_head._unit_loss_grad = False
        if self.world == 1:
            return
        if not self._reduced:
            dist.all_reduce(self.flat_g)
            return
        comm = self._comm()
        comm.wait_stream(torch.cuda.current_stream())
        cur = 0
        with torch.cuda.stream(comm):
            for a, b in sorted(self._reduced) + [(self.flat_g.numel(), self.flat_g.numel())]:
                if a > cur:
                    dist.all_reduce(self.flat_g[cur:a])
                cur = max(cur, b)
        torch.cuda.current_stream().wait_stream(comm)

    _in_step = False

    def lr_at(self, it):
        """mmcv LrUpdaterHook, warmup='linear': lr = base * (1 - (1 - it/warmup_iters) * (1 - ratio))."""
        regular = self.base_lr * self.lr_gamma ** sum(1 for s_ in self.lr_steps if it >= s_)
        if it < self.warmup_iters:
            k = (1 - it / self.warmup_iters) * (1 - self.warmup_ratio)
            return regular * (1 - k)
        return regular

    def state_dict(self):
        """Optimizer state for checkpoint / resume (mmcv CheckpointHook saves `optimizer` + `meta.iter`): momentum
        buffer (flat, same order as `self.params`), iteration counter and hyper-parameters."""
        return dict(iter=self.iter, momentum_buffer=self.flat_m.detach().clone(), base_lr=self.base_lr,
                    momentum=self.momentum, weight_decay=self.wd, warmup_iters=self.warmup_iters,
                    warmup_ratio=self.warmup_ratio, lr_steps=self.lr_steps, lr_gamma=self.lr_gamma,
                    param_shapes=[tuple(p.shape) for p in self.params])

    def load_state_dict(self, sd):
        assert [tuple(p.shape) for p in self.params] == list(sd['param_shapes']), 'optimizer state of another model'
        self.flat_m.copy_(sd['momentum_buffer'])
        self.iter = int(sd['iter'])
        self.base_lr, self.momentum, self.wd = sd['base_lr'], sd['momentum'], sd['weight_decay']
        self.warmup_iters, self.warmup_ratio = sd['warmup_iters'], sd['warmup_ratio']
        self.lr_steps, self.lr_gamma = tuple(sd['lr_steps']), sd['lr_gamma']
        engine.clear_step_cache(self.params)
        if hasattr(self, '_hyper'):
            self._set_hyper()

    def step(self, data, sync_logs=True):
        with engine.precision(self.precision):
            return self._step(data, sync_logs)

    def _step(self, data, sync_logs):
        self._eager_steps = getattr(self, '_eager_steps', 0) + 1
        self.flat_g.zero_()
        engine.zero_pool_begin(self.flat_g.device)
        self._begin_exchange()
        losses = self.model(**data)
        if sync_logs:
            loss, log_vars = self.model._parse_losses(losses)
        else:
            loss = sum(sum(v) for k, v in losses.items() if 'loss' in k)
            log_vars = None
        self._backward(loss)
        engine.zero_pool_end()
        self._finish_exchange()
        ops.sgd_flat(self.flat_w, self.flat_g, self.flat_m, self.lr_at(self.iter), self.momentum, self.wd,
                     1.0 / self.world, first_step=(self.iter == 0))
        engine.clear_step_cache(self.params)
        self.iter += 1
        return loss.detach(), log_vars

    # ------------------------------------------------------------------ CUDA-graph replay
    def _body(self, data):
        """Everything of one iteration that depends only on device-resident state: captured once."""
        with engine.precision(self.precision):
            self._body_inner(data)

    def _body_inner(self, data):
        self.flat_g.zero_()
        engine.zero_pool_begin(self.flat_g.device)
        self._begin_exchange()
        losses = self.model(**data)
        terms = torch.stack([sum(v) for v in losses.values()])
        loss = terms.sum()
        self._backward(loss)
        engine.zero_pool_end()
        self._finish_exchange()
        ops.sgd_flat_dev(self.flat_w, self.flat_g, self.flat_m, self._hyper)
        self._terms.copy_(terms.detach())
        self._loss.copy_(loss.detach())

    def capture(self, data, warmup=2):
        """Capture forward + fused loss + backward + all-reduce + SGD into one CUDA graph.

        Target assignment stays outside the graph (its grid depends on the number of gt boxes): each
        replay first runs `ldb200_ld_targets` eagerly into the static target buffers the graph reads.
        The image batch is copied into a static device buffer; lr / first-step flag live in device memory.

        The `warmup` eager iterations are REAL training iterations on `data` (weights, momentum and `self.iter`
        advance), run on a side stream as graph capture requires.  Call capture() before any eager `step()` of this
        object: autograd binds a parameter's gradient-accumulation node to the stream of its first use, and nodes bound
        to the default stream are not captured (observed: a capture after eager steps replays with garbage
        gradients)."""
        if warmup < 1:
            raise ValueError('capture() needs >= 1 eager warm-up iteration (lazy initialisation must not happen '
                             'inside stream capture)')
        if getattr(self, '_eager_steps', 0) and os.environ.get('LDB200_CAPTURE_AFTER_EAGER') is None:
            raise RuntimeError('capture() after eager step() calls of the same LDTrainStep: the gradient-accumulation '
                               'nodes of the parameters are bound to the stream of those steps and would run outside '
                               'the capture (wrong gradients on replay).  Capture first, or build a fresh model.')
        dev = self.flat_w.device
        head = self.model.bbox_head
        self._graph_mode = True
        self._hyper = torch.zeros(5, dtype=torch.float32, device=dev)
        self._loss = torch.zeros((), dtype=torch.float32, device=dev)
        self._terms = torch.zeros(8, dtype=torch.float32, device=dev)
        self._static_img = data['img'].to(dev).float().contiguous().clone()
        self._metas = data['img_metas']
        sizes = self._featmap_sizes(self._static_img.shape[2:])
        self._static_targets = head.get_targets(sizes, self._metas, data['gt_bboxes'], data['gt_labels'], dev)
        head._static_targets = self._static_targets
        static = dict(img=self._static_img, img_metas=self._metas, gt_bboxes=None, gt_labels=None)
        self._set_hyper()
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(warmup):
                self._body(static)
                engine.clear_step_cache(self.params)
                self.iter += 1
                self._set_hyper()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        engine.clear_step_cache(self.params)
        self._graph = torch.cuda.CUDAGraph()
        l0 = L.launch_count['n']
        # thread_local: the NCCL watchdog thread may query events while this thread captures
        with torch.cuda.graph(self._graph, capture_error_mode='thread_local'):
            self._body(static)
        self.launches_per_step = L.launch_count['n'] - l0   # ld_b200 kernels inside the graph
        engine.clear_step_cache(self.params)
        return self

    def _featmap_sizes(self, hw):
        def half(v):
            return (v + 1) // 2
        h, w = half(half(half(hw[0]))), half(half(half(hw[1])))
        out = []
        for _ in self.model.bbox_head.anchor_generator.strides:
            out.append((h, w))
            h, w = half(h), half(w)
        return out

    def _set_hyper(self):
        vals = torch.tensor([self.lr_at(self.iter), self.momentum, self.wd, 1.0 / self.world,
                             1.0 if self.iter == 0 else 0.0], dtype=torch.float32)
        self._hyper.copy_(vals, non_blocking=True)

    def prefetch(self, data):
        """Start the host->device copy of the NEXT batch's images on a copy stream (pinned host memory), so that it
        overlaps the current replay; `step_graphed(data)` then only does a device-to-device move."""
        dev = self.flat_w.device
        if not hasattr(self, '_copy_stream'):
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._staging_img = torch.empty_like(self._static_img)
            self._staged = None
            self._d2d_done = None
        if self._d2d_done is not None:      # the staging buffer must have been drained by the last step's D2D move
            self._copy_stream.wait_event(self._d2d_done)   # (NOT the whole replay that was enqueued after it)
        with torch.cuda.stream(self._copy_stream):
            self._staging_img.copy_(data['img'], non_blocking=True)
            # target assignment of the NEXT batch too (it only needs the gt boxes): off the critical path of the replay
            self._staged_targets = self.model.bbox_head.get_targets(
                self._featmap_sizes(self._static_img.shape[2:]), data['img_metas'], data['gt_bboxes'], data['gt_labels'],
                dev, static=False)
        self._staged = data

    def step_graphed(self, data=None):
        """Replay the captured iteration on a new batch (same shapes).  Returns the device loss scalar."""
        if data is not None:
            if getattr(self, '_staged', None) is data:
                torch.cuda.current_stream().wait_stream(self._copy_stream)
                self._static_img.copy_(self._staging_img, non_blocking=True)
                tg = self._staged_targets
                tg.flat.record_stream(torch.cuda.current_stream())
                self._static_targets.flat.copy_(tg.flat, non_blocking=True)     # one copy: every field is a view of it
                self._d2d_done = torch.cuda.Event()
                self._d2d_done.record()
                self._staged = self._staged_targets = None
            else:
                self._static_img.copy_(data['img'], non_blocking=True)
                tg = self.model.bbox_head.get_targets(self._featmap_sizes(self._static_img.shape[2:]), data['img_metas'],
                                                      data['gt_bboxes'], data['gt_labels'], self.flat_w.device,
                                                      static=False)
                self._static_targets.flat.copy_(tg.flat, non_blocking=True)
        self._set_hyper()
        self._graph.replay()
        self.iter += 1
        return self._loss

"""GPU: the input side -- `GpuTrainPipeline` (one `ldb200_preprocess_batch` launch: cv2-exact bilinear
Resize, flip, Normalize, Pad, collate) is BIT-identical to the reference's transform classes (sha256 fingerprints in
tests/golden/p_*.pt) and to the CPU restatement; boxes / metas follow `_resize_bboxes` / `bbox_flip`."""
import hashlib
import os

import numpy as np
import pytest
import torch

from oracle import cases, pipeline_ref as pr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('case', cases.PIPELINE_CASES, ids=lambda c: c[0])
def test_gpu_pipeline_is_bit_identical(golden_dir, case):
    from ld_b200.pipeline import GpuTrainPipeline
    name, sizes, seed, flips, img_scale = case
    gold = torch.load(os.path.join(golden_dir, name + '.pt'), weights_only=False)
    imgs, boxes, labels = pr.synth_images(sizes, seed)
    out = GpuTrainPipeline(img_scale=img_scale)(imgs, boxes, labels, flips=flips)
    torch.cuda.synchronize()
    batch = out['img'].cpu()
    ref, _, _ = pr.train_pipeline(imgs, boxes, flips, img_scale)
    assert batch.shape == ref.shape and torch.equal(batch, ref)            # every pixel, padding included
    for i in range(len(imgs)):
        ph, pw = gold['pad_shape'][i][:2]
        chw = np.ascontiguousarray(batch[i, :, :ph, :pw].numpy())
        assert hashlib.sha256(chw.tobytes()).hexdigest() == gold['sha'][i]   # == the reference's transform classes
        m = out['img_metas'][i]
        assert tuple(m['img_shape']) == gold['img_shape'][i] and tuple(m['pad_shape']) == gold['pad_shape'][i]
        assert torch.equal(out['gt_bboxes'][i].cpu(), gold['boxes'][i])
        assert torch.equal(torch.from_numpy(m['scale_factor']), gold['scale_factor'][i])


def test_pipeline_feeds_the_train_step():
    import ld_b200
    from ld_b200 import presets, synth
    from ld_b200.config import ConfigDict
    from ld_b200.pipeline import GpuTrainPipeline
    imgs, boxes, labels = pr.synth_images([(120, 160), (100, 150)], 5)
    data = GpuTrainPipeline(img_scale=(333, 200))(imgs, boxes, labels, flips=[True, False])
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    synth.seeded_init(m, 2)
    synth.seeded_init(m.teacher_model, 3, prefix='teacher.')
    m = m.cuda().train()
    loss, log = ld_b200.LDTrainStep(m, lr=0.001).step(data)
    assert torch.isfinite(loss) and all(np.isfinite(v) for v in log.values())

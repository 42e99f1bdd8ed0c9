"""CPU: mmcv-format checkpoint I/O (ld_b200/checkpoint.py: save / load / publish) -- round trip through
the oracle port (the reference's parameter names), optimizer state, `module.` prefixes, publish_model naming."""
import hashlib
import os

import pytest
import torch

import ld_b200
from ld_b200 import checkpoint as ck, presets, synth
from ld_b200.config import ConfigDict
from oracle import ld_oracle as lo


def test_save_load_publish_round_trip(tmp_path):
    m = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    synth.seeded_init(m, 4)
    ts = ld_b200.LDTrainStep(m, lr=0.01, lr_steps=(100,))
    ts.flat_m.normal_()
    ts.iter = 77
    f = ck.save_checkpoint(m, str(tmp_path / 'epoch_1.pth'), optimizer=ts, meta=dict(epoch=1, iter=77))
    raw = torch.load(f, weights_only=False)
    assert set(raw) == {'meta', 'state_dict', 'optimizer'} and raw['meta']['epoch'] == 1 and 'time' in raw['meta']
    # the reference-named oracle port loads the file as-is, and the values agree
    om = lo.OracleLDDetector(18, 18)
    res = om.load_state_dict(raw['state_dict'], strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    for k, v in m.state_dict().items():
        assert torch.equal(raw['state_dict'][k], v.detach().cpu()), k
        assert raw['state_dict'][k].is_contiguous()
    # load into a fresh model + optimizer (with a DataParallel-style prefix on the way)
    pref = dict(meta=raw['meta'], state_dict={'module.' + k: v for k, v in raw['state_dict'].items()}, optimizer=raw['optimizer'])
    torch.save(pref, tmp_path / 'prefixed.pth')
    m2 = ld_b200.build_detector(ConfigDict(presets.ld_model('r18', teacher_depth=18)))
    ts2 = ld_b200.LDTrainStep(m2, lr=0.5)
    ck.load_checkpoint(m2, str(tmp_path / 'prefixed.pth'), strict=True, optimizer=ts2)
    for (k, a), (_, b) in zip(m.state_dict().items(), m2.state_dict().items()):
        assert torch.equal(a, b), k
    assert ts2.iter == 77 and ts2.base_lr == 0.01 and torch.equal(ts2.flat_m, ts.flat_m) and ts2.lr_steps == (100,)
    with pytest.raises(FileNotFoundError):
        ck.load_checkpoint(m2, str(tmp_path / 'nope.pth'))
    # publish: optimizer dropped, sha256 prefix in the name
    final = ck.publish_model(f, str(tmp_path / 'ld_r18.pth'))
    pub = torch.load(final, weights_only=False)
    assert 'optimizer' not in pub and set(pub) == {'meta', 'state_dict'}
    sha = hashlib.sha256(open(final, 'rb').read()).hexdigest()[:8]
    assert os.path.basename(final) == f'ld_r18-{sha}.pth'

"""CPU: the native network's parameter table (no GPU needed to build the plan
skeleton) lists exactly the reference's state_dict entries, in order, with the
reference's shapes; and the package's constructor init equals the oracle's
(which is pinned bit-for-bit to the reference by tests/golden/student.json)."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from oracle import student_ref
from vpd_b200 import init as vinit
from vpd_b200.rgb import _NativeNet
from vpd_b200._lib import lib, VpdError


def _sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().numpy()).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize('arch,D,use_flow,motion', [('resnet34', 32, True, True),
                                                    ('resnet18', 26, False, False)])
def test_tensor_table_matches_reference_state_dict(arch, D, use_flow, motion):
    torch.manual_seed(1)
    sd = student_ref.init_encoder_state(arch, D, use_flow)
    net = _NativeNet(arch, D, 5 if use_flow else 3, 128, 128, 4, motion)
    table = net.tensor_table()
    enc = [(n, s) for n, a, o, l, s in table if n.startswith('resnet.')]
    assert [n for n, _ in enc] == list(sd.keys())
    for (n, s), (k, v) in zip(enc, sd.items()):
        assert tuple(s) == tuple(v.shape), n
    dec = [(n, s) for n, a, o, l, s in table if n.startswith('decoder.')]
    if motion:
        dsd = student_ref.init_decoder_state(D)
        assert [n[len('decoder.'):] for n, _ in dec] == list(dsd.keys())
        for (n, s), v in zip(dec, dsd.values()):
            assert tuple(s) == tuple(v.shape)
    else:
        assert dec == []
    # arenas: no overlaps, 16-byte aligned tensors, everything inside the arena
    n_params = lib().call('vpd_net_param_count', net.handle)
    spans = []
    for n, a, o, l, s in table:
        if a != 0:
            continue
        numel = 7 * 64 * 64 if l == 2 else int(np.prod(s))
        assert o % 4 == 0 and o + numel <= n_params, n
        spans.append((o, o + numel))
    spans.sort()
    assert all(a[1] <= b[0] for a, b in zip(spans, spans[1:]))
    assert lib().call('vpd_net_workspace_bytes', net.handle) > 0
    net.close()


def test_unsupported_configs_fail_loudly():
    with pytest.raises(VpdError):
        _NativeNet('resnet50', 32, 5, 128, 128, 4, True)
    with pytest.raises(VpdError):
        _NativeNet('resnet34', 32, 5, 100, 128, 4, True)
    with pytest.raises(NotImplementedError):
        vinit.blocks('wide_resnet50_2')


def test_package_init_is_reference_init(golden_dir):
    with open(os.path.join(golden_dir, 'student.json')) as fp:
        meta = json.load(fp)
    torch.manual_seed(meta['init_seed'])
    sd = vinit.encoder_state('resnet34', 32, True)
    dsd = vinit.decoder_state(32)
    assert _sd_hash(sd) == meta['encoder_init_sha256']
    assert _sd_hash(dsd) == meta['decoder_init_sha256']
    torch.manual_seed(5)
    assert _sd_hash(vinit.encoder_state('resnet18', 26, False)) == \
        meta['resnet18_rgb_D26_seed5_sha256']

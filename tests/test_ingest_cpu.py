"""CPU: crop-directory ingest (vpd_b200.ingest) - the packed shard holds exactly the bytes the
reference's loader decodes, and (video, frame) keys map to pool rows."""
import os

import numpy as np
import pytest

from oracle.gen_golden import write_crop_dir
from vpd_b200 import ingest, synth


def test_pack_and_load_shard_roundtrip(tmp_path):
    crops = str(tmp_path / 'crops')
    vids = {}
    for v, n in (('b_vid', 3), ('a_vid', 2)):
        rgb, flow = synth.crops(n, seed=len(v) + n, height=32, width=32)
        write_crop_dir(crops, v, rgb, flow)
        vids[v] = (rgb.numpy(), flow.numpy())
    os.makedirs(os.path.join(crops, 'empty_vid'))
    prefix = str(tmp_path / 'shard')
    index = ingest.pack_crop_dir(crops, prefix, flow_img='flow', img_dim=32)
    assert [v['name'] for v in index['videos']] == ['a_vid', 'b_vid', 'empty_vid']
    sh = ingest.load_shard(prefix)
    assert len(sh) == 5 and sh.rgb.dtype == np.uint8 and sh.rgb.shape == (5, 32, 32, 3)
    # PNG is lossless: the pools hold the synthetic crops bit for bit (RGB order restored)
    assert np.array_equal(np.asarray(sh.rgb[0:2]), vids['a_vid'][0])
    assert np.array_equal(np.asarray(sh.rgb[2:5]), vids['b_vid'][0])
    assert np.array_equal(np.asarray(sh.flow[2:5]), vids['b_vid'][1])
    rows = sh.rows_of([('b_vid', 2, None), ('a_vid', 0, None), ('b_vid', 0, None)])
    assert rows.tolist() == [4, 0, 2]
    with pytest.raises(KeyError):
        sh.rows_of([('a_vid', 7)])
    vl = sh.videos()
    assert [(v[0], v[1]) for v in vl] == [('a_vid', [0, 1]), ('b_vid', [0, 1, 2]),
                                          ('empty_vid', [])]
    assert vl[1][2].shape == (3, 32, 32, 3)


def test_pack_without_flow_and_resize(tmp_path):
    crops = str(tmp_path / 'crops')
    rgb, flow = synth.crops(2, seed=9, height=48, width=48)
    write_crop_dir(crops, 'v', rgb, flow)
    prefix = str(tmp_path / 's')
    ingest.pack_crop_dir(crops, prefix, flow_img=None, img_dim=32)      # resized like the loader
    sh = ingest.load_shard(prefix)
    assert sh.flow is None and sh.mask is None and sh.rgb.shape == (2, 32, 32, 3)
    assert not os.path.exists(prefix + '.flow.npy')


def test_pack_masks(tmp_path):
    import cv2
    crops = str(tmp_path / 'crops')
    rgb, flow = synth.crops(3, seed=4, height=32, width=32)
    write_crop_dir(crops, 'v', rgb, flow)
    m = np.zeros((32, 32, 3), np.uint8)
    m[4:9, :, :] = 255
    cv2.imwrite(os.path.join(crops, 'v', '1.mask.png'), m)              # only frame 1 has a mask
    prefix = str(tmp_path / 's')
    ingest.pack_crop_dir(crops, prefix, flow_img='flow', img_dim=32, with_mask=True)
    sh = ingest.load_shard(prefix)
    assert sh.mask.shape == (3, 32, 32) and sh.mask.dtype == np.uint8
    assert sh.mask[0].max() == 0 and sh.mask[2].max() == 0
    assert np.array_equal(np.asarray(sh.mask[1]), m[:, :, 0])

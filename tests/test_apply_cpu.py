"""CPU: host logic of the apply driver - video sharding (incl. a 2-rank gloo run) and the
pickle format of apply_vpd_model.py:163-178."""
import os
import pickle
import subprocess
import sys

import numpy as np

from vpd_b200 import apply as vapply

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_videos_partitions_and_balances():
    counts = [4200, 10, 3900, 4100, 0, 2500, 2600, 700]
    for world in (1, 2, 3, 8):
        parts = [vapply.shard_videos(counts, world, r) for r in range(world)]
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(len(counts)))
        loads = [sum(counts[i] for i in p) for p in parts]
        if world == 2:
            assert max(loads) - min(loads) <= max(counts)
    assert vapply.shard_videos([], 4, 1) == []


def test_plan_chunks_tiles_the_frames_in_order():
    for counts, bs in (([5, 0, 12, 3], 8), ([2695] * 3, 500), ([1], 500), ([], 4), ([4, 4], 4)):
        chunks = vapply.plan_chunks(counts, bs)
        assert all(sum(hi - lo for _, lo, hi, _ in c) == bs for c in chunks[:-1])
        seen = [[] for _ in counts]
        for c in chunks:
            fill = 0
            for v, lo, hi, off in c:
                assert off == fill and 0 <= lo < hi <= counts[v]
                fill += hi - lo
                seen[v].extend(range(lo, hi))
            assert 0 < fill <= bs
        assert seen == [list(range(n)) for n in counts]


def test_pickle_writer_processes(tmp_path):
    w = vapply._PickleWriters(2)
    embs = np.random.RandomState(0).randn(6, 2, 4).astype(np.float32)
    for i in range(5):
        w.submit(os.path.join(str(tmp_path), 'v{}.emb.pkl'.format(i)), [3, 1, 2], embs[i:i + 3] if i < 4 else embs[:3], True)
    w.close()
    with open(os.path.join(str(tmp_path), 'v1.emb.pkl'), 'rb') as fp:
        back = pickle.load(fp)
    assert [t[0] for t in back] == [1, 2, 3] and np.array_equal(back[0][1], embs[2])
    bad = vapply._PickleWriters(1)
    bad.submit(os.path.join(str(tmp_path), 'missing_dir', 'x.pkl'), [0], embs[:1], True)
    try:
        bad.close()
        raise AssertionError('a failed write must surface')
    except FileNotFoundError:
        pass


def test_pickle_format_matches_reference(tmp_path):
    embs = np.arange(3 * 2 * 4, dtype=np.float32).reshape(3, 2, 4)
    out = vapply.format_video_embs([7, 2, 5], embs, flip=True)
    assert [t[0] for t in out] == [2, 5, 7]
    assert all(isinstance(t[0], int) and t[1].dtype == np.float32 and t[1].shape == (2, 4)
               and t[2] == {} for t in out)
    assert np.array_equal(out[0][1], embs[1])
    single = vapply.format_video_embs([1, 0], embs[:2], flip=False)
    assert single[0][1].shape == (4,) and np.array_equal(single[0][1], embs[1, 0])
    path = os.path.join(str(tmp_path), 'v.emb.pkl')
    vapply.store_pickle(path, out)
    with open(path, 'rb') as fp:
        back = pickle.load(fp)
    assert back[2][0] == 7 and np.array_equal(back[2][1], embs[0])


GLOO_SCRIPT = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from vpd_b200 import apply as vapply
dist.init_process_group('gloo')
rank, world = dist.get_rank(), dist.get_world_size()
counts = [50, 7, 31, 44, 12, 9]
mine = vapply.shard_videos(counts, world, rank)
# gradient-sum semantics used by the trainer: all_reduce(SUM) of a flat arena
g = torch.full((5,), float(rank + 1))
dist.all_reduce(g, op=dist.ReduceOp.SUM)
assert g.tolist() == [3.0] * 5
gathered = [None] * world
dist.all_gather_object(gathered, mine)
if rank == 0:
    flat = sorted(i for p in gathered for i in p)
    assert flat == list(range(len(counts))), gathered
    print('OK', gathered)
dist.destroy_process_group()
'''


def test_two_rank_gloo_sharding(tmp_path):
    script = os.path.join(str(tmp_path), 'gloo_shard.py')
    with open(script, 'w') as fp:
        fp.write(GLOO_SCRIPT)
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', '29611', script, ROOT]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=240)
    assert res.returncode == 0, res.stderr[-2000:]
    assert 'OK' in res.stdout


def test_downstream_reference_loader_reads_our_pickles(tmp_path):
    """the reference's downstream loader (action_dataset/load.py:16-64, used by
    recognize.py) must consume the `.emb.pkl` files this package writes. Runs only where the
    reference checkout is present (the build container)."""
    import sys
    import pytest
    from oracle import ref_shim
    if not ref_shim.available():
        pytest.skip('reference checkout not present')
    sys.dont_write_bytecode = True
    if ref_shim.REFERENCE_DIR not in sys.path:
        sys.path.insert(0, ref_shim.REFERENCE_DIR)
    from action_dataset.load import load_embs
    rng = np.random.RandomState(0)
    embs = rng.randn(4, 2, 32).astype(np.float32)
    frames = [3, 0, 1, 6]                                 # gaps: the loader interpolates
    vapply.store_pickle(os.path.join(str(tmp_path), 'clipA.emb.pkl'),
                        vapply.format_video_embs(frames, embs, flip=True))
    vapply.store_pickle(os.path.join(str(tmp_path), 'clipB.emb.pkl'),
                        vapply.format_video_embs([0, 1], embs[:2], flip=True))
    out = load_embs(str(tmp_path), norm=False)
    assert sorted(out) == ['clipA', 'clipB']
    dense, mask = out['clipA']
    assert dense.shape == (7, 2, 32) and mask.tolist() == [True, True, False, True, False, False,
                                                           True]
    assert np.allclose(dense[3], embs[0]) and np.allclose(dense[6], embs[3])
    assert np.allclose(dense[2], 0.5 * dense[1] + 0.5 * dense[3])   # its gap interpolation
    normed = load_embs(str(tmp_path), norm=True)['clipB'][0]
    assert np.allclose(np.linalg.norm(normed, axis=2), 1.0)


def test_tennis_crop_layout(tmp_path):
    """apply_vpd_model.get_tennis_dataset's layout: <crop_dir>/<src_video>/<player>/<abs frame>.png,
    one output video per player and clip, frame numbers relative to the clip start"""
    import cv2
    import torch
    crop_dir = str(tmp_path)
    g = torch.Generator().manual_seed(3)
    imgs = {}
    for player, frames in (('front', [100, 101, 103]), ('back', [])):
        d = os.path.join(crop_dir, 'match_a', player)
        os.makedirs(d)
        for f in frames:
            rgb = torch.randint(0, 256, (16, 16, 3), generator=g, dtype=torch.uint8).numpy()
            flow = torch.randint(0, 256, (16, 16, 3), generator=g, dtype=torch.uint8).numpy()
            cv2.imwrite(os.path.join(d, '{}.png'.format(f)), cv2.cvtColor(rgb, cv2.COLOR_RGB2BGR))
            cv2.imwrite(os.path.join(d, '{}.flow.png'.format(f)), flow)
            imgs[f] = (rgb, flow)
    videos = vapply.read_tennis_crops(crop_dir, ['match_a_100_104'], flow_img='flow', img_dim=16)
    assert [v[0] for v in videos] == ['front__match_a_100_104', 'back__match_a_100_104']
    name, frames, rgb, flow = videos[0]
    assert frames == [0, 1, 3] and rgb.shape == (3, 16, 16, 3) and flow.shape == (3, 16, 16, 3)
    assert np.array_equal(rgb[2].numpy(), imgs[103][0]) and np.array_equal(flow[0].numpy(), imgs[100][1])
    assert videos[1][1] == [] and videos[1][2].shape[0] == 0
    # the generic reader sees the same files under their directory name
    gen = vapply.read_crop_dir(os.path.join(crop_dir, 'match_a'), flow_img='flow', img_dim=16)
    assert [v[0] for v in gen] == ['back', 'front'] and gen[1][1] == [100, 101, 103]
    assert np.array_equal(gen[1][2].numpy(), rgb.numpy())

#!/usr/bin/env python3
"""Generate tests/golden/* by running the UNMODIFIED reference on the CPU.

Run in the build container only (needs /root/reference):
    python -m oracle.gen_golden
Everything written is small (< 200 KB total) and committed. The inputs are
regenerated from seeds by the tests (vpd_b200/synth.py), so only reference
OUTPUTS (and hashes of large ones) are stored.
"""
import hashlib
import json
import os
import random
import shutil
import sys
import tempfile

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim                      # noqa: E402
from vpd_b200 import synth                       # noqa: E402

GOLD = os.path.join(ROOT, 'tests', 'golden')


def sha(t):
    if isinstance(t, torch.Tensor):
        t = t.detach().contiguous().numpy()
    return hashlib.sha256(np.ascontiguousarray(t).tobytes()).hexdigest()


def sd_hash(sd):
    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(np.ascontiguousarray(v.detach().numpy()).tobytes())
    return h.hexdigest()


def write_crop_dir(root, video, rgb, flow, flow_name='flow'):
    import cv2
    d = os.path.join(root, video)
    os.makedirs(d, exist_ok=True)
    for i in range(rgb.shape[0]):
        cv2.imwrite(os.path.join(d, '{}.png'.format(i)),
                    cv2.cvtColor(rgb[i].numpy(), cv2.COLOR_RGB2BGR))
        cv2.imwrite(os.path.join(d, '{}.{}.png'.format(i, flow_name)), flow[i].numpy())


def gen_assembly(ref):
    """A1-A4: FrameDataset / GenericDataset on PNGs of synthetic crops."""
    out = {}
    tmp = tempfile.mkdtemp()
    try:
        # -- per-dataset LUTs through the real _load_image/_load_flow ------
        ramp = torch.arange(256, dtype=torch.uint8).view(16, 16, 1).repeat(1, 1, 3)
        write_crop_dir(tmp, 'ramp', ramp[None], ramp[None])
        luts = {}
        for name, ms in ref.RGB_MEAN_STD.items():
            ds = ref.FrameDataset([(0, 0, os.path.join(tmp, 'ramp', '0'))], 16, ms,
                                  flow_img_name='flow')
            img = ds[0]['img'][0].numpy()                 # [5,16,16]
            luts[name] = img.reshape(5, 256)
        out['lut_names'] = np.array(sorted(luts))
        out['luts'] = np.stack([luts[k] for k in sorted(luts)])

        # -- apply path, 32x32, flip on / off ------------------------------
        rgb, flow = synth.crops(4, seed=11, height=32, width=32)
        write_crop_dir(tmp, 'v32', rgb, flow)
        ms = ref.RGB_MEAN_STD['fs']
        tasks = [(0, i, os.path.join(tmp, 'v32', str(i))) for i in range(4)]
        ds = ref.FrameDataset(tasks, 32, ms, augment_flip=True, flow_img_name='flow')
        out['apply_flip'] = torch.stack([ds[i]['img'] for i in range(4)]).numpy()
        ds = ref.FrameDataset(tasks, 32, ms, augment_flip=False, flow_img_name='flow')
        out['apply_noflip'] = torch.stack([ds[i]['img'] for i in range(4)]).numpy()
        ds = ref.FrameDataset(tasks, 32, ms, augment_flip=True, flow_img_name=None)
        out['apply_flip_rgbonly'] = torch.stack([ds[i]['img'] for i in range(4)]).numpy()

        # -- train path: GenericDataset.__getitem__ with the stochastic
        #    augmentations neutralised (no jitter in the transform, identity
        #    crop, no mask png) and the flip bit forced --------------------
        teach = synth.teacher(4, seed=12, emb_dim=8, motion=True).numpy()
        data = [('v32', i, teach[i], {}) for i in range(4)]
        gd = ref.GenericDataset(data, tmp, 32, ms, 4, augment=False,
                                flow_img_name='flow')
        gd.augment = True                      # enables the flip branch only:
        gd._random_crop = lambda im: im        # transform was built w/o jitter
        flips = [0, 1, 1, 0]
        imgs, embs = [], []
        sf = ref.single_frame
        orig_choice, orig_bits = random.choice, random.getrandbits
        try:
            for i in range(4):
                random.choice = lambda seq, i=i: seq[i]
                random.getrandbits = lambda k, i=i: flips[i]
                item = gd[0]
                imgs.append(item['img'])
                embs.append(item['emb'])
        finally:
            random.choice, random.getrandbits = orig_choice, orig_bits
        out['train_flips'] = np.array(flips, dtype=np.uint8)
        out['train_img'] = torch.stack(imgs).numpy()
        out['train_emb'] = torch.stack(embs).numpy()

        # -- one full-size frame, hash only ---------------------------------
        rgb, flow = synth.crops(1, seed=13)
        write_crop_dir(tmp, 'v128', rgb, flow)
        ds = ref.FrameDataset([(0, 0, os.path.join(tmp, 'v128', '0'))], 128, ms,
                              augment_flip=True, flow_img_name='flow')
        out['apply128_sha256'] = np.array(sha(ds[0]['img']))
    finally:
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(GOLD, 'assembly.npz'), **out)
    print('assembly.npz written')


def gen_student(ref):
    """A5-A10: constructor, embed, train steps, loss curve."""
    meta = {}
    D = 32
    torch.manual_seed(0)
    enc = ref.RGBF_EmbeddingModel('resnet34', D, True, 'cpu')
    trainer = ref.ModelTrainer(enc, True)
    meta['init_seed'] = 0
    meta['encoder_init_sha256'] = sd_hash(enc.state_dict())
    meta['decoder_init_sha256'] = sd_hash(trainer.fcn_time.state_dict())
    torch.manual_seed(5)
    enc18 = ref.RGBF_EmbeddingModel('resnet18', 26, False, 'cpu')
    meta['resnet18_rgb_D26_seed5_sha256'] = sd_hash(enc18.state_dict())

    # eval-mode embed with non-trivial BN buffers
    from oracle.student_ref import randomize_bn_state
    sd = randomize_bn_state(enc.state_dict(), seed=21)
    enc.load_state_dict(sd)
    rgb, flow = synth.crops(4, seed=22)
    from oracle import assemble_ref
    x = assemble_ref.apply_batch(rgb.numpy(), flow.numpy(), *synth.FS_MEAN_STD,
                                 flip=True).view(-1, 5, 128, 128)
    emb = enc.embed(x)
    arrays = {'embed_out': emb}

    # two training steps through ModelTrainer.epoch (B=8), losses + checksums
    torch.manual_seed(0)
    enc = ref.RGBF_EmbeddingModel('resnet34', D, True, 'cpu')
    trainer = ref.ModelTrainer(enc, True)
    opt, scaler = trainer.get_optimizer(5e-4)
    assert scaler is None
    B, steps = 8, 200
    rgb, flow = synth.crops(64, seed=1)
    teach = synth.teacher(64, seed=3, emb_dim=D, motion=True)
    fl = synth.flips(steps * B, seed=2)
    idx_g = torch.Generator().manual_seed(4)
    idx = torch.randint(0, 64, (steps * B,), generator=idx_g)
    losses = []
    for s in range(steps):
        sel = idx[s * B:(s + 1) * B]
        f = fl[s * B:(s + 1) * B]
        img, tgt = assemble_ref.train_batch(
            rgb[sel].numpy(), flow[sel].numpy(), teach[sel].numpy(), f.numpy(),
            *synth.FS_MEAN_STD)
        if s == 0:
            # gradients of the very first step, before the optimizer touches anything
            enc.train(); trainer.fcn_time.train()
            import copy
            enc2 = copy.deepcopy(enc); dec2 = copy.deepcopy(trainer.fcn_time)
            out = dec2(enc2(img))
            l0 = torch.nn.functional.mse_loss(out, tgt, reduction='sum')
            l0.backward()
            names = [n for n, _ in enc2.named_parameters()] + \
                    ['decoder.' + n for n, _ in dec2.named_parameters()]
            params = list(enc2.parameters()) + list(dec2.parameters())
            arrays['step0_out'] = out.detach().numpy()
            arrays['step0_grad_norms'] = np.array(
                [p.grad.norm().item() for p in params], dtype=np.float64)
            arrays['step0_grad_fc'] = enc2.resnet.fc.weight.grad.numpy()
            arrays['step0_grad_conv1'] = enc2.resnet.conv1.weight.grad.numpy()
            arrays['step0_grad_l4'] = enc2.resnet.layer4[2].conv2.weight.grad[:8, :8].numpy()
            meta['param_names'] = names
        loss = trainer.epoch([{'img': img, 'emb': tgt}], optimizer=opt, scaler=scaler)
        losses.append(loss)
        if s in (0, 1):
            meta['step{}_state_sha256'.format(s)] = sd_hash(enc.state_dict())
            arrays['step{}_fc_weight'.format(s)] = enc.resnet.fc.weight.detach().numpy().copy()
            arrays['step{}_bn1_running_var'.format(s)] = enc.resnet.bn1.running_var.numpy().copy()
        if s % 20 == 0:
            print('step', s, 'loss/frame', loss, flush=True)
    arrays['loss_curve'] = np.array(losses, dtype=np.float64)
    meta['loss_curve'] = {'batch': B, 'steps': steps, 'lr': 5e-4, 'pool': 64,
                          'seeds': {'crops': 1, 'flips': 2, 'teacher': 3, 'index': 4}}
    # eval-mode loss on a fixed batch after training (exercises running stats)
    sel = torch.arange(8)
    img, tgt = assemble_ref.train_batch(rgb[sel].numpy(), flow[sel].numpy(),
                                        teach[sel].numpy(), np.zeros(8, np.uint8),
                                        *synth.FS_MEAN_STD)
    meta['final_eval_loss'] = trainer.epoch([{'img': img, 'emb': tgt}])
    np.savez_compressed(os.path.join(GOLD, 'student.npz'), **arrays)
    with open(os.path.join(GOLD, 'student.json'), 'w') as fp:
        json.dump(meta, fp, indent=1)
    print('student.npz / student.json written')


def write_teacher_pickles(root):
    """Synthetic teacher output: 3 videos, [2, 8] embeddings, gaps in the frame numbers and
    low-score frames so that every branch of load_default is exercised."""
    import pickle
    rng = np.random.RandomState(21)
    for v, frames in (('vidA', [0, 1, 2, 3, 5, 6, 7, 9]), ('vidB', [10, 11, 12, 13, 14]),
                      ('skip_me', [0, 1, 2])):
        embs = []
        for j, f in enumerate(frames):
            meta = {'dp_score': 0.9} if j % 3 else {'kp_score': 0.3 if f in (2, 12) else 0.8}
            embs.append((f, rng.randn(2, 8).astype(np.float32), meta))
        with open(os.path.join(root, v + '.emb.pkl'), 'wb') as fp:
            pickle.dump(embs, fp)
    with open(os.path.join(root, 'notes.txt'), 'w') as fp:
        fp.write('not a pickle')


def write_tennis_pickles(root):
    """`<player>__<video>_<start>_<end>.emb.pkl` as the tennis teacher run writes them"""
    import pickle
    rng = np.random.RandomState(22)
    for stem, frames in (('front__match_a_100_140', [0, 1, 2, 3, 4, 6, 7]),
                         ('back__match_a_100_140', [0, 1, 2, 5, 6]),
                         ('front__match_b_set_2_7_30', [3, 4, 5, 6])):
        embs = []
        for j, f in enumerate(frames):
            meta = {'kp_score': 0.3 if j == 2 else 0.9}
            embs.append((f, rng.randn(2, 8).astype(np.float32), meta))
        with open(os.path.join(root, stem + '.emb.pkl'), 'wb') as fp:
            pickle.dump(embs, fp)


def gen_targets(ref):
    """A13: GenericDataset.load_default on synthetic teacher pickles."""
    tmp = tempfile.mkdtemp()
    arrays, meta = {}, {}
    real_listdir = os.listdir
    os.listdir = lambda d: sorted(real_listdir(d))   # directory order is file-system dependent
    try:
        write_teacher_pickles(tmp)
        for name, kw in (('motion', dict(embed_time=True)),
                         ('plain_norm', dict(embed_time=False, normalize_target=True)),
                         ('motion_norm_excl', dict(embed_time=True, normalize_target=True,
                                                   min_pose_score=0.2,
                                                   exclude_prefixes=('skip',)))):
            np.random.seed(5)
            tr, va, D = ref.GenericDataset.load_default(
                tmp, tmp, 128, kw.pop('embed_time'), 100, ([0.5] * 3, [0.2] * 3), **kw)
            for part, ds in (('train', tr), ('val', va)):
                meta['{}_{}_keys'.format(name, part)] = [[d[0], int(d[1])] for d in ds.data]
                arrays['{}_{}'.format(name, part)] = np.stack([d[2] for d in ds.data])
            meta[name + '_emb_dim'] = int(D)
        # TennisDataset.load_default: per player-and-clip pickles, frame = clip start + index
        tmp2 = os.path.join(tmp, 'tennis')
        os.makedirs(tmp2)
        write_tennis_pickles(tmp2)
        for name, kw in (('tennis_motion', dict(embed_time=True)),
                         ('tennis_plain', dict(embed_time=False, min_pose_score=0.2))):
            np.random.seed(6)
            tr, va, D = ref.single_frame.TennisDataset.load_default(
                tmp2, tmp2, 128, kw.pop('embed_time'), 100, ([0.5] * 3, [0.2] * 3), **kw)
            for part, ds in (('train', tr), ('val', va)):
                meta['{}_{}_keys'.format(name, part)] = [[d[0], d[1], int(d[2])] for d in ds.data]
                arrays['{}_{}'.format(name, part)] = np.stack([d[3] for d in ds.data])
            meta[name + '_emb_dim'] = int(D)
    finally:
        os.listdir = real_listdir
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(GOLD, 'targets.npz'), **arrays)
    with open(os.path.join(GOLD, 'targets.json'), 'w') as fp:
        json.dump(meta, fp, indent=1)
    print('targets.npz / targets.json written')


def gen_augment(ref):
    """A3 with augment=True: the real GenericDataset.__getitem__ (ColorJitter, masked noise,
    flip, RandomResizedCrop) after `random.seed(s); torch.manual_seed(s)`. Only the outputs are
    stored; tests re-draw the parameters with vpd_b200.augment.draw_batch from the same seeds."""
    import cv2
    from oracle import augment_ref
    out = {}
    tmp = tempfile.mkdtemp()
    try:
        ms = ref.RGB_MEAN_STD['fs']
        for tag, n, dim, items, seed in (('s32', 6, 32, 12, 21), ('s128', 3, 128, 3, 22)):
            rgb, flow, mask, has_mask = augment_ref.augment_inputs(n, seed, dim, dim)
            video = 'aug' + tag
            write_crop_dir(tmp, video, rgb, flow)
            for i in range(n):
                if bool(has_mask[i]):
                    cv2.imwrite(os.path.join(tmp, video, '{}.mask.png'.format(i)),
                                mask[i].numpy()[:, :, None].repeat(3, axis=2))
            teach = synth.teacher(n, seed=seed + 1, emb_dim=8, motion=True).numpy()
            data = [(video, i, teach[i], {}) for i in range(n)]
            gd = ref.GenericDataset(data, tmp, dim, ms, items, augment=True, flow_img_name='flow')
            random.seed(seed)
            torch.manual_seed(seed)
            imgs, embs = [], []
            for k in range(items):
                item = gd[k]
                imgs.append(item['img'])
                embs.append(item['emb'])
            img = torch.stack(imgs).numpy()
            out[tag + '_emb'] = torch.stack(embs).numpy()
            out[tag + '_seed'] = np.array(seed)
            if dim > 32:                       # keep the fixture small: every 4th pixel + a hash
                out[tag + '_img_sub4'] = img[:, :, ::4, ::4].copy()
            else:
                out[tag + '_img'] = img
    finally:
        shutil.rmtree(tmp)
    np.savez_compressed(os.path.join(GOLD, 'augment.npz'), **out)
    print('augment.npz written')


def gen_keypoint(ref):
    """VIPE* teacher apply path: the reference's FCResNet + Keypoint_EmbeddingModel.embed and
    apply_vipe_model.mean_embs_by_frame on seeded inputs."""
    import types
    from oracle import keypoint_ref
    sys.modules.setdefault('matplotlib', types.ModuleType('matplotlib'))
    from models.module import FCResNet
    from models.keypoint import Keypoint_EmbeddingModel
    import apply_vipe_model
    out = {}
    for tag, in_dim, joints, hidden, blocks, seed in (('d39', 39, 13, 1024, 2, 5),
                                                     ('d75', 75, 25, 256, 1, 6)):
        torch.manual_seed(seed)
        enc = FCResNet(in_dim, 32, blocks, hidden, dropout=0.2)
        out[tag + '_init_sha256'] = np.array(sd_hash(enc.state_dict()))
        enc.load_state_dict(keypoint_ref.perturb_bn(enc.state_dict(), seed + 100))
        model = Keypoint_EmbeddingModel(enc, {}, 'cpu')
        poses = keypoint_ref.synth_poses(96, seed + 200, joints)
        out[tag + '_emb'] = model.embed(poses)
        out[tag + '_emb_one'] = model.embed(poses[3].numpy())           # [J,3] -> [1,D]
        out[tag + '_seed'] = np.array(seed)
    # mean_embs_by_frame: 2 detections in some frames, flipped twins
    g = torch.Generator().manual_seed(9)
    embs = []
    for frame in (7, 3, 3, 11):
        for fl in (False, True):
            embs.append((frame, torch.randn(32, generator=g).numpy(),
                         {'kp_score': float(torch.rand(1, generator=g)), 'is_mean': False,
                          'is_flip': fl}))
    # normalize_2d_skeleton on synthetic COCO detections (pixel coordinates), all four variants
    from vipe_dataset.dataset_base import normalize_2d_skeleton
    g2 = torch.Generator().manual_seed(10)
    kp = torch.rand((6, 17, 3), generator=g2)
    kp[:, :, :2] = kp[:, :, :2] * 300 + 50
    kp[5, [5, 6, 11, 12], :2] = 77.0                     # degenerate torso: zero extent
    kp = kp.numpy()
    out['skel_in'] = kp
    for fl in (False, True):
        for bones in (False, True):
            out['skel_f{}_b{}'.format(int(fl), int(bones))] = np.stack([
                normalize_2d_skeleton(kp[i], fl, include_bone_features=bones).numpy()
                for i in range(6)])
    res = apply_vipe_model.mean_embs_by_frame(embs, True)
    out['mean_frames'] = np.array([r[0] for r in res])
    out['mean_embs'] = np.stack([r[1] for r in res])
    out['mean_scores'] = np.array([r[2]['kp_score'] for r in res])
    out['mean_is_mean'] = np.array([r[2]['is_mean'] for r in res])
    np.savez_compressed(os.path.join(GOLD, 'keypoint.npz'), **out)
    print('keypoint.npz written')


def gen_keypoint_train(ref):
    """VIPE* teacher training: two optimizer steps of the reference's
    Keypoint_EmbeddingModel.epoch (two datasets zipped: one with negatives + 3-D targets, one
    with pose pairs only) - losses, first-step gradients, BatchNorm buffers, parameters after
    the steps; the dropout masks the reference drew (replayed, checked through the oracle)."""
    import types
    from oracle import keypoint_train_ref as T
    sys.modules.setdefault('matplotlib', types.ModuleType('matplotlib'))
    from models.module import FCResNet, FCPoseDecoder
    from models.keypoint import Keypoint_EmbeddingModel
    H, blocks, n1, n2, p, lr = 128, 2, 136, 72, 0.2, 1e-3
    torch.manual_seed(31)
    enc = FCResNet(39, 32, blocks, H, dropout=p)
    dec = FCPoseDecoder(32, [128, 128], [('h36m', 140)], dropout=0)
    out = {'enc_init_sha256': np.array(sd_hash(enc.state_dict())),
           'dec_init_sha256': np.array(sd_hash(dec.state_dict()))}
    model = Keypoint_EmbeddingModel(enc, {'3d': dec}, 'cpu')
    params = list(enc.parameters()) + list(dec.parameters())
    names = (['enc.' + k for k, _ in enc.named_parameters()]
             + ['dec.' + k for k, _ in dec.named_parameters()])
    real = torch.optim.AdamW(params, lr=lr)
    captured = []

    class Opt:                                   # records the gradients the reference computed
        def step(self):
            captured.append({k: v.grad.clone() for k, v in zip(names, params)})
            real.step()

        def zero_grad(self):
            real.zero_grad()

    fcn_keys = ['fcn.layers.0', 'fcn.layers.2']
    for step_i in range(2):
        b1 = T.synth_batch(n1, 40 + step_i)
        b2 = T.synth_batch(n2, 50 + step_i, with_neg=False, with_3d=False)
        enc_sd0 = {k: v.clone() for k, v in enc.state_dict().items()}
        dec_sd0 = {k: v.clone() for k, v in dec.state_dict().items()}
        torch.manual_seed(60 + step_i)
        contra, loss, per = model.epoch([('h36m', [b1]), ('pair', [b2])], optimizer=Opt(),
                                        weight_3d=1)
        # replay the masks the reference drew and check the restatement against it
        torch.manual_seed(60 + step_i)
        masks = [T.replay_masks(n1, H, blocks, 3, p), T.replay_masks(n2, H, blocks, 2, p)]
        res, grads, bufs = T.zipped_step(enc_sd0, dec_sd0, fcn_keys, [('h36m', b1), ('pair', b2)],
                                         masks, p, blocks)
        tot = sum(v[1] for v in res.values()) / (n1 + n2)
        assert abs(tot - loss) <= 1e-5 * abs(loss), (tot, loss)
        for k, g in captured[step_i].items():
            assert torch.allclose(g, grads[k], rtol=1e-4, atol=1e-7), k
        for k, v in bufs.items():
            assert torch.allclose(v.float(), enc.state_dict()[k].float(), rtol=1e-5, atol=1e-6), k
        out['step{}_contra'.format(step_i)] = np.array(contra)
        out['step{}_loss'.format(step_i)] = np.array(loss)
        out['step{}_loss_h36m'.format(step_i)] = np.array(per['h36m'])
        out['step{}_loss_pair'.format(step_i)] = np.array(per['pair'])
        for di, m in enumerate(masks):
            packed = np.packbits(np.stack([torch.stack(q).numpy() for q in m]).astype(np.uint8))
            out['step{}_masks{}'.format(step_i, di)] = packed
    for k in ('enc.layers.0.weight', 'enc.layers.0.bias', 'enc.layers.2.block.0.weight',
              'enc.layers.2.block.1.weight', 'enc.layers.2.block.1.bias',
              'enc.layers.3.block.4.weight', 'enc.layers.3.block.5.weight', 'enc.layers.4.weight',
              'enc.layers.4.bias', 'dec.fcn.layers.0.weight', 'dec.fcn.layers.2.bias',
              'dec.fc_h36m.weight', 'dec.fc_h36m.bias'):
        out['grad0/' + k] = captured[0][k].numpy()
    for k in ('layers.2.block.1.running_mean', 'layers.2.block.1.running_var',
              'layers.3.block.5.running_mean', 'layers.3.block.5.running_var',
              'layers.2.block.1.num_batches_tracked'):
        out['final/' + k] = enc.state_dict()[k].numpy()
    out['final/enc.layers.4.weight'] = enc.state_dict()['layers.4.weight'].numpy()
    out['final/dec.fc_h36m.bias'] = dec.state_dict()['fc_h36m.bias'].numpy()
    # evaluation epoch after the two steps (running statistics, no dropout)
    contra, loss, per = model.epoch([('h36m', [T.synth_batch(n1, 70)])])
    out['eval_contra'], out['eval_loss'] = np.array(contra), np.array(loss)
    np.savez_compressed(os.path.join(GOLD, 'keypoint_train.npz'), **out)
    print('keypoint_train.npz written')


def main():
    os.makedirs(GOLD, exist_ok=True)
    ref = ref_shim.load()
    torch.set_num_threads(os.cpu_count())
    which = sys.argv[1:] or ['assembly', 'student', 'targets', 'augment', 'keypoint', 'keypoint_train']
    if 'assembly' in which:
        gen_assembly(ref)
    if 'student' in which:
        gen_student(ref)
    if 'targets' in which:
        gen_targets(ref)
    if 'augment' in which:
        gen_augment(ref)
    if 'keypoint' in which:
        gen_keypoint(ref)
    if 'keypoint_train' in which:
        gen_keypoint_train(ref)
    with open(os.path.join(GOLD, 'README.md'), 'w') as fp:
        fp.write('Golden vectors produced by `python -m oracle.gen_golden` from the unmodified\n'
                 'reference at /root/reference (torch {}, CPU fp32). Inputs are regenerated\n'
                 'from seeds by vpd_b200/synth.py; see oracle/gen_golden.py.\n'.format(
                     torch.__version__))


if __name__ == '__main__':
    main()

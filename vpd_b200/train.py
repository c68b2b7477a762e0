"""Training driver for the student path: the epoch loop of `train_vpd_model.main`
(reference train_vpd_model.py:214-283) around `ModelTrainer`, plus a loader that assembles
batches on the GPU from device-resident uint8 crop pools (K1) instead of PNG-decoding
DataLoader workers.

    fit(trainer, train_loader, val_loader, save_dir, config, num_epochs, ...)

writes exactly what the reference writes: `config.json` (keys train_vpd_model.py:222-228),
`loss.json` (one record per epoch, :251-262), `best_epoch.{encoder,decoder}.pt` whenever the
`model_select_window`-epoch moving average of the validation loss improves (:257-269),
`epochNNNN.*` every `checkpoint_frequency` epochs and for the last epoch (:270-280).
`trainer` is anything with the ModelTrainer interface (epoch / save_model), so the loop is
testable without a GPU.
"""
import json
import os

import numpy as np

CONFIG_KEYS = ('num_epochs', 'batch_size', 'learning_rate', 'img_dim', 'use_flow', 'motion',
               'emb_dim', 'encoder_arch', 'rgb_mean_std')


def get_moving_avg_loss(losses, n, key):      # train_vpd_model.py:114-115
    return np.mean([l[key] for l in losses[-n:]])


def fit(trainer, train_loader, val_loader, save_dir, config, num_epochs, optimizer, scaler=None,
        model_select_window=5, checkpoint_frequency=None, dataset='synthetic', log=print):
    """Runs `num_epochs` epochs; returns the loss history (the content of loss.json)."""
    from . import dp
    missing = [k for k in CONFIG_KEYS if k not in config]
    assert not missing, 'config lacks {}'.format(missing)
    # data parallel: rank 0 owns the run directory (config, loss history, checkpoints - the BN
    # running statistics saved are rank 0's); the epoch losses are reduced over all ranks, so
    # every rank takes the same branches below
    main_rank = dp.rank() == 0
    if main_rank:
        os.makedirs(save_dir)                  # like the reference: refuses to overwrite a run
        with open(os.path.join(save_dir, 'config.json'), 'w') as fp:
            json.dump({k: config[k] for k in CONFIG_KEYS}, fp, indent=2)
    else:
        log = lambda *a, **k: None             # noqa: E731
    dp.barrier()
    loss_file = os.path.join(save_dir, 'loss.json')
    losses = []
    best_val_loss = float('inf')
    epoch = 0
    for epoch in range(1, num_epochs + 1):
        train_loss = trainer.epoch(train_loader, optimizer=optimizer, scaler=scaler)
        val_loss = float('nan')
        if val_loader is not None:
            val_loss = trainer.epoch(val_loader)
        losses.append({'epoch': epoch, 'train': train_loss, 'val': val_loss,
                       'dataset_train': [(dataset, train_loss)],
                       'dataset_val': [(dataset, val_loss)]})
        moving_avg_val_loss = get_moving_avg_loss(losses, model_select_window, 'val')
        log('Epoch {} - train loss: {:0.4f} [avg: {:0.4f}] val loss: {:0.4f} [avg: {:0.4f}]'.format(
            epoch, train_loss, get_moving_avg_loss(losses, model_select_window, 'train'),
            val_loss, moving_avg_val_loss))
        if main_rank:
            with open(loss_file, 'w') as fp:
                json.dump(losses, fp, indent=2)
        if moving_avg_val_loss < best_val_loss:
            log('New best epoch!')
            if main_rank:
                trainer.save_model(save_dir, 'best_epoch')
        if checkpoint_frequency is not None and epoch % checkpoint_frequency == 0:
            log('Saving checkpoint: {}'.format(epoch))
            if main_rank:
                trainer.save_model(save_dir, 'epoch{:04d}'.format(epoch))
        best_val_loss = min(moving_avg_val_loss, best_val_loss)
    if epoch > 0:
        log('Saving last epoch: {}'.format(epoch))
        if main_rank:
            trainer.save_model(save_dir, 'epoch{:04d}'.format(epoch))
    dp.barrier()
    return losses


class PoolLoader:
    """Epoch of `target_len` frames drawn WITH replacement (the reference's `_TrainDataset`:
    `__len__` = target_len, `_get` = random.choice, vpd_dataset/common.py:104-108) from
    device-resident uint8 pools; every batch is assembled by the K1 kernel (normalisation,
    random horizontal flip with the flow-x sign change, teacher row = emb[int(flip)]).
    Yields the reference's batch dict {'img': f32 [B,5,H,W], 'emb': f32 [B,Dt]} on the GPU.
    With `mask_u8` [P,H,W] (first channel of `<n>.mask.png`, zeros where a frame has none) the
    masked-noise augmentation of single_frame.py:179-191 is applied on the device with its
    per-frame coin (p = 0.5).
    `augment=True` is the reference's `augment=True` dataset in full (single_frame.py:168-206):
    ColorJitter, masked noise, flip and RandomResizedCrop, drawn per frame in the reference's
    order from Python `random` and torch's global generator (vpd_b200/augment.py - seed THOSE
    to reproduce a single-process reference loader; `seed` is not used then) and applied by
    the K1a kernel; `has_mask` [P] says which frames have a mask PNG (default: all, if
    `mask_u8` is given); `host_noise=True` also draws the noise on the host like the reference
    (exact, but 3*H*W floats per noisy frame over PCIe) instead of the device generator.
    `fast_draws=True` draws the same distributions in batched calls from the loader's own
    generator (`seed`): 0.3 ms instead of 22 ms of host time per 256 frames, not
    stream-identical with the reference.
    Flips follow the reference (single_frame.py:171-174): a frame is flipped only when the
    dataset augments AND its teacher entry has the two rows [as is, flipped]; with
    `augment=False` no frame is flipped unless `random_flip=True` asks for the flip alone
    (still only with two-row teachers - a one-row target has no flipped counterpart).
    `raw=True` (plain loader only: no augmentation, no mask noise) yields the draw itself -
    {'rgb_u8': pool, 'flow_u8': pool, 'index': int32 [B], 'flip', 'teacher': pool,
    'rgb_mean_std'} - which `ModelTrainer.epoch` assembles straight into the network's input
    layout (one K1 launch, 23 us per 256 frames, the same bf16 values) instead of an fp32
    reference-format batch plus a layout conversion."""

    def __init__(self, rgb_u8, flow_u8, teacher, rgb_mean_std, batch_size, target_len, seed=0,
                 mask_u8=None, augment=False, has_mask=None, host_noise=False, fast_draws=False,
                 random_flip=False, raw=False):
        import torch
        self.rgb, self.flow, self.teacher = rgb_u8, flow_u8, teacher
        self.mask = mask_u8
        self.augment, self.has_mask, self.host_noise = augment, has_mask, host_noise
        self.fast_draws = fast_draws
        self.random_flip = random_flip
        self.raw = bool(raw) and not augment and mask_u8 is None and teacher is not None
        if augment and mask_u8 is not None and has_mask is None:
            self.has_mask = torch.ones(rgb_u8.shape[0], dtype=torch.bool)
        self.rgb_mean_std = rgb_mean_std
        self.batch_size, self.target_len = batch_size, target_len
        self.gen = torch.Generator().manual_seed(seed)

    def __len__(self):
        return (self.target_len + self.batch_size - 1) // self.batch_size

    def __iter__(self):
        import torch
        from .assemble import assemble_batch
        dev = self.rgb.device
        n = self.rgb.shape[0]
        left = self.target_len
        while left > 0:
            b = min(self.batch_size, left)
            left -= b
            if self.augment:
                from .assemble import assemble_batch_aug
                from .augment import draw_batch
                two_rows = self.teacher is not None and self.teacher.dim() == 3
                if self.fast_draws:      # same distributions, batched draws from self.gen
                    from .augment import draw_batch_fast
                    p = draw_batch_fast(b, n, self.rgb.shape[1], self.rgb.shape[2],
                                        two_rows=two_rows, has_mask=self.has_mask,
                                        generator=self.gen).to(dev)
                else:
                    p = draw_batch(b, n, self.rgb.shape[1], self.rgb.shape[2], two_rows=two_rows,
                                   has_mask=self.has_mask, host_noise=self.host_noise,
                                   channels=5 if self.flow is not None else 3).to(dev)
                yield assemble_batch_aug(self.rgb, self.flow, self.rgb_mean_std, p,
                                         teacher=self.teacher, mask=self.mask)
                continue
            idx = torch.randint(0, n, (b,), generator=self.gen).int().to(dev)
            flip = None
            if self.random_flip and self.teacher is not None and self.teacher.dim() == 3:
                flip = torch.randint(0, 2, (b,), generator=self.gen).to(torch.uint8).to(dev)
            if self.raw:
                yield {'rgb_u8': self.rgb, 'flow_u8': self.flow, 'index': idx, 'flip': flip,
                       'teacher': self.teacher, 'rgb_mean_std': self.rgb_mean_std}
                continue
            kw = {}
            if self.mask is not None:
                from .assemble import RANDOM_MASK_PROB
                coin = (torch.rand((b,), generator=self.gen) <= RANDOM_MASK_PROB)
                kw = dict(mask=self.mask, noise_on=coin.to(torch.uint8).to(dev),
                          seed=int(torch.randint(0, 2 ** 62, (1,), generator=self.gen).item()))
            yield assemble_batch(self.rgb, self.flow, self.rgb_mean_std, flip=flip,
                                 teacher=self.teacher, index=idx, **kw)


def main(emb_dir, shard_prefix, save_dir, rgb_mean_std, dataset='generic', num_epochs=50,
         batch_size=256, learning_rate=5e-4, img_dim=128, motion=False, encoder_arch='resnet34',
         model_select_window=5, checkpoint_frequency=25, min_pose_score=None,
         exclude_prefixes=None, augment=True, target_len=20000, device='cuda', log=print,
         _factories=None):
    """`train_vpd_model.main` (:171-283) on this package, from the packed crop shard
    (`vpd_b200.ingest.pack_crop_dir`) and the teacher pickles: targets (A13; `dataset='tennis'`
    selects the per player-and-clip variant, :119-128) -> 80/20 split -> device-resident pools ->
    `PoolLoader`s of `target_len` / 20 % of it draws per epoch (:183, single_frame.py:268-272;
    the reference augments the validation set too, common.py:86) -> model, trainer, AdamW ->
    `fit`. Returns the loss history. `_factories` (tests) replaces the GPU-side constructors:
    {'pools': fn(shard, rows) -> (rgb, flow, mask), 'loader': fn(...), 'trainer': fn(emb_dim,
    use_flow) -> (trainer, optimizer, scaler)}."""
    import torch
    from . import targets
    from .ingest import load_shard
    f = _factories or {}
    if dataset == 'tennis':
        data, emb_dim = targets.load_teacher_targets_tennis(
            emb_dir, motion, min_pose_score=min_pose_score, exclude_prefixes=exclude_prefixes)
        train_data, val_data = targets.split_train_val(data, key_len=3)
        key = lambda d: ('{}/{}'.format(d[0], d[1]), d[2])        # shard video = <video>/<player>
    else:
        data, emb_dim = targets.load_teacher_targets(
            emb_dir, motion, min_pose_score=min_pose_score, exclude_prefixes=exclude_prefixes)
        train_data, val_data = targets.split_train_val(data)
        key = lambda d: (d[0], d[1])
    shard = load_shard(shard_prefix)
    use_flow = shard.flow is not None
    log('Train images: {}  Val images: {}  Embedding dim: {}'.format(
        len(train_data), len(val_data), emb_dim))

    def pools(part):
        rows = shard.rows_of([key(d) for d in part])
        if 'pools' in f:
            return f['pools'](shard, rows)
        # numpy fancy indexing on the memory-mapped shard reads just these rows into a fresh
        # (writable) array
        rgb = torch.from_numpy(np.asarray(shard.rgb)[rows]).to(device)
        flow = torch.from_numpy(np.asarray(shard.flow)[rows]).to(device) if use_flow else None
        mask = None
        if shard.mask is not None:
            mask = torch.from_numpy(np.asarray(shard.mask)[rows]).to(device)
        return rgb, flow, mask

    def loader(part, length):
        rgb, flow, mask = pools(part)
        teach = targets.targets_array(part)
        if 'loader' in f:
            return f['loader'](rgb, flow, mask, teach, length)
        return PoolLoader(rgb, flow, torch.from_numpy(teach).to(device), rgb_mean_std, batch_size,
                          length, mask_u8=mask, augment=augment, raw=True)

    train_loader = loader(train_data, target_len)
    val_loader = loader(val_data, int(target_len * 0.2)) if val_data else None
    if 'trainer' in f:
        trainer, optimizer, scaler = f['trainer'](emb_dim, use_flow)
    else:
        from .rgb import RGBF_EmbeddingModel
        from .trainer import ModelTrainer
        trainer = ModelTrainer(RGBF_EmbeddingModel(encoder_arch, emb_dim, use_flow, device), motion)
        optimizer, scaler = trainer.get_optimizer(learning_rate)
    config = {'num_epochs': num_epochs, 'batch_size': batch_size, 'learning_rate': learning_rate,
              'img_dim': img_dim, 'use_flow': use_flow, 'motion': motion, 'emb_dim': emb_dim,
              'encoder_arch': encoder_arch,
              'rgb_mean_std': [list(map(float, rgb_mean_std[0])), list(map(float, rgb_mean_std[1]))]}
    return fit(trainer, train_loader, val_loader, save_dir, config, num_epochs, optimizer,
               scaler=scaler, model_select_window=model_select_window,
               checkpoint_frequency=checkpoint_frequency, dataset=dataset, log=log)

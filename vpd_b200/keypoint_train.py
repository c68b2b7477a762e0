"""Training step of the keypoint (VIPE*) teacher on the B200 (BASELINE config 4).

Reference: `Keypoint_EmbeddingModel.epoch` models/keypoint.py:38-126 (per dataset batch: three
weight-sharing encoder passes pose1 / pose2 / pose_neg, hinge losses on the embedding distances,
3-D pose decoder + sum-MSE on `kp_features`, losses summed over the datasets of a zipped batch,
divided by the sample count, backward, optimizer step), `FCResNet` / `FcResidualBlock`
models/module.py:159-204 in train mode, `FCPoseDecoder` :230-246, `step` models/util.py:50-58,
AdamW over encoder + decoder parameters train_vipe_model.py:164-169,312-314.

Everything is composed from C-ABI launches (no torch arithmetic on the data path):
  Linear layers      vpd_conv2d_fwd / _dgrad / _wgrad as 1x1 convolutions over [n,1,1,C]
                     (bf16 operands, fp32 accumulation; BatchNorm column statistics come out of
                     the forward epilogue in fp64)
  BN1d+ReLU+Dropout  vpd_bn1d_fwd / vpd_bn1d_bwd (+ the block's `x2 - x`)
  last encoder Linear vpd_linear_rows_f32 forward (fp32 embedding), conv dgrad / wgrad backward
  loss head          vpd_vipe_loss
  optimizer          vpd_adamw over one flat fp32 arena (padded rows / columns stay zero)
Parameters live in that arena; `state_dict()` entries are views into it.

Deliberate, output-neutral differences from the reference:
  * a Linear bias in front of a batch-statistics BatchNorm has an exactly-zero gradient in
    exact arithmetic (the reference's autograd produces rounding noise there); we write zero.
    The bias still enters `running_mean`.
  * dropout masks come from the device generator (vpd_dropout_mask) unless `masks=` supplies
    them (tests replay the reference's CPU draws that way).
"""
from collections import OrderedDict

import torch

from . import init as _init
from ._lib import lib, stream_ptr, VpdError


def _pad64(v):
    return (v + 63) // 64 * 64


class _Arena:
    """flat fp32 parameter / gradient storage with named, possibly padded, 2-D / 1-D entries"""

    def __init__(self):
        self.entries = OrderedDict()       # name -> (offset, padded shape, logical shape)
        self.size = 0

    def add(self, name, shape, padded=None):
        padded = tuple(padded or shape)
        n = 1
        for d in padded:
            n *= d
        self.entries[name] = (self.size, padded, tuple(shape))
        self.size += (n + 3) // 4 * 4      # keep every entry 16-byte aligned

    def alloc(self, dev):
        self.params = torch.zeros(self.size, device=dev, dtype=torch.float32)
        self.grads = torch.zeros(self.size, device=dev, dtype=torch.float32)

    def full(self, name, grad=False):
        off, padded, _ = self.entries[name]
        n = 1
        for d in padded:
            n *= d
        return (self.grads if grad else self.params)[off:off + n].view(padded)

    def view(self, name, grad=False):
        """the logical (unpadded) tensor as a view"""
        _, _, shape = self.entries[name]
        t = self.full(name, grad)
        return t[tuple(slice(0, s) for s in shape)]


class FCPoseDecoder:
    """models/module.py:230-246 parameter container (state_dict keys / init as the reference):
    FCNet(emb_dim, hidden_dims[:-1], hidden_dims[-1]) + one Linear per target."""

    def __init__(self, emb_dim, hidden_dims, target_dims, dropout=0):
        assert len(hidden_dims) >= 2
        if dropout != 0:
            raise NotImplementedError('decoder dropout (the reference trains with 0)')
        if any(h % 64 for h in hidden_dims):
            raise NotImplementedError('CUDA path: decoder hidden dims must be multiples of 64')
        self.emb_dim, self.hidden_dims = emb_dim, list(hidden_dims)
        self.target_dims = list(target_dims)
        self.training = True
        sd = OrderedDict()
        dims = [emb_dim] + list(hidden_dims)
        # FCNet.__init__ numbering (models/module.py:138-152): Linear at 0, then per hidden
        # layer ReLU, Linear and - between hidden layers only - Dropout
        self.fcn_keys = ['fcn.layers.0']
        last = 0
        for i in range(len(hidden_dims) - 1):
            lin = last + 2
            self.fcn_keys.append('fcn.layers.{}'.format(lin))
            last = lin + (1 if i + 1 < len(hidden_dims) - 1 else 0)
        for i, key in enumerate(self.fcn_keys):
            sd[key + '.weight'], sd[key + '.bias'] = _init._draw_linear(dims[i + 1], dims[i])
        for name, tdim in self.target_dims:
            key = 'fc_{}'.format(name)
            sd[key + '.weight'], sd[key + '.bias'] = _init._draw_linear(tdim, hidden_dims[-1])
        self._sd = sd

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd, strict=True):
        if strict and set(sd) != set(self._sd):
            raise RuntimeError('FCPoseDecoder: state_dict keys differ')
        for k in self._sd:
            if k in sd:
                self._sd[k].copy_(sd[k].detach().to(self._sd[k].device))      # keeps arena views
        self._version = getattr(self, '_version', 0) + 1    # KeypointTrainCore re-packs its mirrors

    def parameters(self):
        return list(self._sd.values())

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def to(self, device):
        return self


class KeypointAdamW:
    """torch.optim.AdamW defaults over the trainer's flat arena (train_vipe_model.py:312-314)"""

    def __init__(self, core, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.core, self.lr, self.betas, self.eps, self.weight_decay = core, lr, betas, eps, weight_decay
        self.step_count = 0
        self._m = torch.zeros_like(core.arena.params)
        self._v = torch.zeros_like(core.arena.params)

    def step(self):
        a = self.core.arena
        self.step_count += 1
        lib().call('vpd_adamw', a.params, a.grads, self._m, self._v, a.params.numel(), self.lr,
                   self.betas[0], self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0,
                   stream_ptr(self.core.dev))
        self.core.weights_dirty = True
        self.core._refresh_mirrors()       # bf16 mirrors follow the masters right away

    def zero_grad(self, set_to_none=False):
        self.core.arena.grads.zero_()

    def state_dict(self):
        return {'step': self.step_count, 'exp_avg': self._m.cpu(), 'exp_avg_sq': self._v.cpu(),
                'lr': self.lr, 'betas': self.betas, 'eps': self.eps,
                'weight_decay': self.weight_decay,
                'layout': [(k, v[0], list(v[1])) for k, v in self.core.arena.entries.items()]}

    def load_state_dict(self, sd):
        self._m.copy_(sd['exp_avg'])
        self._v.copy_(sd['exp_avg_sq'])
        self.step_count = sd['step']


class KeypointTrainCore:
    """Owns the arena, the bf16 weight mirrors and the launch sequence of one training step."""

    def __init__(self, encoder, decoder, dev):
        self.enc, self.dec, self.dev = encoder, decoder, dev
        H, D = encoder.hidden_dim, encoder.out_dim
        self.H, self.D, self.cin = H, D, encoder._cin
        self.p_drop = float(encoder.dropout)
        a = _Arena()
        a.add('enc.layers.0.weight', (H, encoder.in_dim), (H, self.cin))
        a.add('enc.layers.0.bias', (H,))
        self.bn_names = []
        for i in range(encoder.num_blocks):
            p = 'enc.layers.{}.block'.format(2 + i)
            for lin, bn in ((0, 1), (4, 5)):
                a.add('{}.{}.weight'.format(p, lin), (H, H))
                a.add('{}.{}.bias'.format(p, lin), (H,))
                a.add('{}.{}.weight'.format(p, bn), (H,))
                a.add('{}.{}.bias'.format(p, bn), (H,))
                self.bn_names.append('{}.{}'.format(p, bn))
        last = 'enc.layers.{}'.format(2 + encoder.num_blocks)
        a.add(last + '.weight', (D, H), (64, H))
        a.add(last + '.bias', (D,), (64,))
        self.last = last
        if decoder is not None:
            dims = [D] + decoder.hidden_dims
            for i, key in enumerate(decoder.fcn_keys):
                a.add('dec.' + key + '.weight', (dims[i + 1], dims[i]), (dims[i + 1], _pad64(dims[i])))
                a.add('dec.' + key + '.bias', (dims[i + 1],))
            for name, tdim in decoder.target_dims:
                a.add('dec.fc_{}.weight'.format(name), (tdim, dims[-1]), (_pad64(tdim), dims[-1]))
                a.add('dec.fc_{}.bias'.format(name), (tdim,), (_pad64(tdim),))
        a.alloc(dev)
        self.arena = a
        # move the current parameter values in; state_dict entries become arena views
        for mod, prefix in ((encoder, 'enc.'), (decoder, 'dec.')):
            if mod is None:
                continue
            for k in list(mod._sd):
                if prefix + k in a.entries:
                    a.view(prefix + k).copy_(mod._sd[k].to(dev))
                    mod._sd[k] = a.view(prefix + k)
                else:
                    mod._sd[k] = mod._sd[k].to(dev)
        encoder._prepared = None
        self.weights_dirty = True
        self._mirrors = {}
        self._ones = {}
        self.loss_sums = torch.zeros(2, device=dev, dtype=torch.float64)
        self.mask_seed = 0x5eed
        self.mask_calls = 0
        # device-side step counter added to the dropout seed: a captured step draws fresh masks
        # on every replay
        self._seed_step = torch.zeros(1, device=dev, dtype=torch.int64)
        self._graphs = {}
        # measured on B200 (tests/diag_keypoint.py, n = 4096): 3.08 ms/step launched one by one,
        # 3.26 ms replayed - the step is bound by the latency of ~160 small kernels, not by
        # launching them; replay only frees the host thread, so it is opt-in
        self.use_graphs = False

    # ---- helpers ---------------------------------------------------------------------------
    def _st(self):
        return stream_ptr(self.dev)

    def _bf(self, *shape):
        return torch.empty(shape, device=self.dev, dtype=torch.bfloat16)

    def _one(self, n):
        if n not in self._ones:
            self._ones[n] = torch.ones(n, device=self.dev, dtype=torch.float32)
        return self._ones[n]

    def _refresh_mirrors(self):
        """bf16 tiled copies of every Linear weight (forward and transposed), once per step"""
        # a load_state_dict on either module (it copies into the arena views in place) bumps
        # the module's version: the mirrors then no longer match the masters
        versions = tuple(getattr(m, '_version', 0) for m in (self.enc, self.dec) if m is not None)
        if versions != getattr(self, '_seen_versions', None):
            self._seen_versions = versions
            self.weights_dirty = True
        if not self.weights_dirty:
            return
        L, st = lib(), self._st()
        for name, (_, padded, _) in self.arena.entries.items():
            if not name.endswith('.weight') or len(padded) != 2:
                continue
            cout, cin = padded
            if name not in self._mirrors:
                self._mirrors[name] = (self._bf(cout * cin), self._bf(cout * cin))
            w, wt = self._mirrors[name]
            L.call('vpd_pack_conv_weight', self.arena.full(name), w, wt, cout, cin, 1, st)
        self.weights_dirty = False

    def _fwd(self, x, name, y, relu, bias=True, stats=None):
        """y = [relu](x . W^T [+ b]) as a 1x1 convolution; stats: vpd_stat_acc [2][Cout] (+=, no bias)"""
        cout, cin = self.arena.entries[name + '.weight'][1]
        n = x.shape[0]
        scale = shift = None
        if bias and stats is None:
            scale, shift = self._one(cout), self.arena.full(name + '.bias')
        lib().call('vpd_conv2d_fwd', x, self._mirrors[name + '.weight'][0], y, n, 1, 1, cin, cout, 1,
                   1, 0, scale, shift, None, int(relu), stats, self._st())

    def _bwd_linear(self, x, dy, name, dx=None, residual=None, bias=True):
        """dW += dy^T x, db += colsum(dy), dx = dy . W (+ residual)"""
        L, st = lib(), self._st()
        cout, cin = self.arena.entries[name + '.weight'][1]
        n = x.shape[0]
        L.call('vpd_conv2d_wgrad', x, dy, self.arena.full(name + '.weight', grad=True), n, 1, 1, cin,
               cout, 1, 1, 0, st)
        if bias:
            L.call('vpd_colsum_bf16', dy, self.arena.full(name + '.bias', grad=True), n, cout, st)
        if dx is not None:
            L.call('vpd_conv2d_dgrad', dy, self._mirrors[name + '.weight'][1], dx, n, 1, 1, cin, cout,
                   1, 1, 0, residual, None, None, 0, st)

    def _mask(self, n, given):
        if given is not None:
            return given
        keep = torch.empty((n, self.H), device=self.dev, dtype=torch.uint8)
        self.mask_calls += 1
        lib().call('vpd_dropout_mask', keep, n * self.H, self.p_drop, self.mask_seed,
                   self._seed_step, self.mask_calls & 0x7fffffff, self._st())
        return keep

    # ---- encoder ---------------------------------------------------------------------------
    def encoder_forward(self, pose, groups=1, masks=None):
        """train-mode pass over `groups` stacked pose batches: pose fp32 [groups * n, in_dim]
        (the weight-sharing passes pose1 / pose2 / pose_neg of one step share every launch;
        each block of n rows is its own BatchNorm batch) -> (emb fp32 [groups * n, D], ctx).
        masks: 2 * num_blocks uint8 [groups * n, H] keep masks, or None (device generator)."""
        L, st, H, enc = lib(), self._st(), self.H, self.enc
        N = pose.shape[0]
        n = N // groups
        ctx = {'n': n, 'N': N, 'groups': groups, 'blocks': []}
        xb = self._bf(N, self.cin)
        L.call('vpd_rows_to_bf16', pose.contiguous(), xb, N, enc.in_dim, self.cin, st)
        h = self._bf(N, H)
        self._fwd(xb, 'enc.layers.0', h, relu=True)
        ctx['xb'], ctx['h0'] = xb, h
        sd = enc._sd
        for i in range(enc.num_blocks):
            p = 'layers.{}.block'.format(2 + i)
            blk = {'h_in': h, 'stage': []}
            z = h
            for j, (lin, bn) in enumerate(((0, 1), (4, 5))):
                a = self._bf(N, H)
                self._fwd(z, 'enc.{}.{}'.format(p, lin), a, relu=False, bias=False)
                stats = torch.empty((groups * 2 * H, 2), device=self.dev, dtype=torch.int64)
                L.call('vpd_colstats_bf16', a, stats, n, H, groups, st)
                keep = self._mask(N, None if masks is None else masks[2 * i + j])
                out = self._bf(N, H)
                b = '{}.{}'.format(p, bn)
                save_mean = torch.empty(groups * H, device=self.dev)
                save_rstd = torch.empty(groups * H, device=self.dev)
                L.call('vpd_bn1d_fwd', a, stats, sd[b + '.weight'], sd[b + '.bias'],
                       sd['{}.{}.bias'.format(p, lin)], sd[b + '.running_mean'],
                       sd[b + '.running_var'], sd[b + '.num_batches_tracked'], save_mean, save_rstd,
                       keep, self.p_drop, h if j == 1 else None, out, n, H, groups, st)
                blk['stage'].append({'x': z, 'a': a, 'keep': keep, 'mean': save_mean,
                                     'rstd': save_rstd, 'lin': 'enc.{}.{}'.format(p, lin),
                                     'bn': 'enc.' + b})
                z = out
            h = z
            ctx['blocks'].append(blk)
        emb = torch.empty((N, self.D), device=self.dev, dtype=torch.float32)
        L.call('vpd_linear_rows_f32', h, self.arena.view(self.last + '.weight'),
               self.arena.view(self.last + '.bias'), emb, N, H, self.D, st)
        ctx['h_last'] = h
        return emb, ctx

    def encoder_backward(self, ctx, de_bf16):
        """de_bf16: bf16 [groups * n, 64] gradient of the embeddings (columns >= D zero)"""
        L, st, H, n, N, G = lib(), self._st(), self.H, ctx['n'], ctx['N'], ctx['groups']
        dh = self._bf(N, H)
        self._bwd_linear(ctx['h_last'], de_bf16, self.last, dx=dh)
        sums = torch.empty((G * 2 * H, 2), device=self.dev, dtype=torch.int64)
        for blk in reversed(ctx['blocks']):
            s2, s1 = blk['stage'][1], blk['stage'][0]
            da2 = self._bf(N, H)
            self._bn_bwd(dh, s2, sums, da2, n, G)
            dz1 = self._bf(N, H)
            self._bwd_linear(s2['x'], da2, s2['lin'], dx=dz1, bias=False)
            da1 = self._bf(N, H)
            self._bn_bwd(dz1, s1, sums, da1, n, G)
            neg = self._bf(N, H)
            L.call('vpd_axpby_bf16', dh, -1.0, None, 0.0, neg, N * H, st)       # d(x2 - x)/dx
            dh_in = self._bf(N, H)
            self._bwd_linear(s1['x'], da1, s1['lin'], dx=dh_in, residual=neg, bias=False)
            dh = dh_in
        dpre = self._bf(N, H)
        L.call('vpd_relu_mask_bf16', dh, ctx['h0'], dpre, N * H, st)
        self._bwd_linear(ctx['xb'], dpre, 'enc.layers.0')

    def _bn_bwd(self, dz, s, sums, da, n, groups):
        g = lambda k: self.arena.full(s['bn'] + k)
        lib().call('vpd_bn1d_bwd', dz, s['a'], s['keep'], self.p_drop, g('.weight'), g('.bias'),
                   s['mean'], s['rstd'], sums, da, self.arena.full(s['bn'] + '.weight', grad=True),
                   self.arena.full(s['bn'] + '.bias', grad=True), n, self.H, groups, self._st())

    # ---- decoder ---------------------------------------------------------------------------
    def decoder_forward(self, emb, target):
        L, st, dec = lib(), self._st(), self.dec
        n = emb.shape[0]
        eb = self._bf(n, 64)
        L.call('vpd_rows_to_bf16', emb, eb, n, self.D, 64, st)
        acts = [eb]
        for key in dec.fcn_keys:
            cout = self.arena.entries['dec.' + key + '.weight'][1][0]
            y = self._bf(n, cout)
            self._fwd(acts[-1], 'dec.' + key, y, relu=True)
            acts.append(y)
        name = 'dec.fc_{}'.format(target)
        tpad = self.arena.entries[name + '.weight'][1][0]
        pred = self._bf(n, tpad)
        self._fwd(acts[-1], name, pred, relu=False)
        return pred, {'acts': acts, 'head': name, 'n': n}

    def decoder_backward(self, ctx, dpred, de_contra_bf16, out=None):
        """-> bf16 [n, 64] (written into `out` if given): decoder gradient of the embedding +
        de_contra"""
        L, st, n = lib(), self._st(), ctx['n']
        acts = ctx['acts']
        d = self._bf(n, acts[-1].shape[1])
        self._bwd_linear(acts[-1], dpred, ctx['head'], dx=d)
        for i in range(len(self.dec.fcn_keys) - 1, -1, -1):
            dm = self._bf(*d.shape)
            L.call('vpd_relu_mask_bf16', d, acts[i + 1], dm, dm.numel(), st)
            dx = out if (i == 0 and out is not None) else self._bf(n, acts[i].shape[1])
            self._bwd_linear(acts[i], dm, 'dec.' + self.dec.fcn_keys[i], dx=dx,
                             residual=de_contra_bf16 if i == 0 else None)
            d = dx
        return d

    # ---- one dataset batch: forward, loss, backward (gradients accumulate) -------------------
    def dataset_step(self, batch, target, gscale, weight_3d=1.0, masks=None, train=True):
        """batch: dict of device fp32 tensors 'pose1' [n, in] (+ 'pose2', 'pose_neg',
        'pose_neg_is_valid' [n], 'kp_features' [n, T]); gscale = 1 / (samples in the zipped
        batch). Adds (contra, loss) to self.loss_sums; returns n.

        With `use_graphs` (opt-in) a training step with device-generated dropout is captured,
        per (dataset, shapes, scale), into a CUDA graph over static input buffers on its second
        call and replayed afterwards (~160 launches -> one; fresh masks come from the
        device-side seed counter)."""
        self._refresh_mirrors()
        n = batch['pose1'].shape[0]
        if not (train and masks is None and self.use_graphs):
            return self._dataset_step(batch, target, gscale, weight_3d, masks, train)
        self._seed_step.add_(1)
        key = (target, float(gscale), float(weight_3d),
               tuple((k, tuple(v.shape)) for k, v in sorted(batch.items())))
        ent = self._graphs.get(key)
        if ent is None:                       # first time: plain launches (also warms every kernel)
            self._graphs[key] = {'calls': 1}
            return self._dataset_step(batch, target, gscale, weight_3d, None, True)
        if 'graph' not in ent:
            ent['static'] = {k: v.clone() for k, v in batch.items()}
            self.mask_calls = 0
            torch.cuda.synchronize(self.dev)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._dataset_step(ent['static'], target, gscale, weight_3d, None, True)
            ent['graph'] = g
        for k, v in batch.items():
            ent['static'][k].copy_(v)
        ent['graph'].replay()
        return n

    def _dataset_step(self, batch, target, gscale, weight_3d, masks, train):
        L, st, D = lib(), self._st(), self.D
        n = batch['pose1'].shape[0]
        names = [k for k in ('pose1', 'pose2', 'pose_neg') if k in batch]
        G = len(names)
        # the weight-sharing passes are stacked along the rows: one launch per layer for all
        pose = batch[names[0]] if G == 1 else torch.cat([batch[k] for k in names])
        if train:
            m = None
            if masks is not None:      # [pass][layer] -> [layer] stacked like the poses
                m = [torch.cat([masks[g][l] for g in range(G)]).contiguous()
                     for l in range(2 * self.enc.num_blocks)]
            emb, ctx = self.encoder_forward(pose, G, m)
        else:                          # evaluation: running statistics, no dropout
            emb = self.enc.eval().forward(pose)
        rows = {k: slice(i * n, (i + 1) * n) for i, k in enumerate(names)}
        has3d = 'kp_features' in batch
        nd = (1 + ('pose2' in rows)) if has3d else 0            # decoder: pose1 (+ pose2) rows
        pred = dctx = None
        if has3d:
            pred, dctx = self.decoder_forward(emb[:nd * n], target)
        de = torch.empty((G * n, D), device=self.dev, dtype=torch.float32)
        dpred = torch.empty_like(pred) if has3d else None
        true3d = batch['kp_features'].reshape(n, -1).contiguous() if has3d else None
        T = true3d.shape[1] if has3d else 0
        tpad = pred.shape[1] if has3d else 0
        part = lambda t, k: t[rows[k]] if (t is not None and k in rows) else None
        two = has3d and nd == 2
        L.call('vpd_vipe_loss', emb[rows['pose1']], part(emb, 'pose2'), part(emb, 'pose_neg'),
               batch.get('pose_neg_is_valid'), pred[:n] if has3d else None,
               pred[n:2 * n] if two else None, true3d, de[rows['pose1']], part(de, 'pose2'),
               part(de, 'pose_neg'), dpred[:n] if has3d else None, dpred[n:2 * n] if two else None,
               self.loss_sums, n, D, T, tpad, float(weight_3d), float(gscale), st)
        if not train:
            return n
        db = self._bf(G * n, 64)
        L.call('vpd_rows_to_bf16', de, db, G * n, D, 64, st)
        if has3d:
            tot = self._bf(G * n, 64)
            self.decoder_backward(dctx, dpred, db[:nd * n], out=tot[:nd * n])
            if nd < G:
                tot[nd * n:].copy_(db[nd * n:])
            db = tot
        self.encoder_backward(ctx, db)
        return n


# ---- the training driver: train_vipe_model.py:171-209,228-229,312-430 ---------------------------
CONFIG_KEYS = ('datasets', 'num_epochs', 'learning_rate', 'batch_size', 'embedding_dim',
               'encoder_arch', 'decoder_arch', 'embed_bones', 'augment_camera')


def save_model(save_dir, name, encoder, decoders, optimizer):
    """train_vipe_model.py:171-183: `<name>.encoder.pt`, `<name>.decoder-<k>.pt`,
    `<name>.optimizer.pt` (the optimizer file holds the fused AdamW's flat moments, not
    torch.optim.AdamW's per-tensor state)"""
    import os
    torch.save(OrderedDict((k, v.cpu()) for k, v in encoder.state_dict().items()),
               os.path.join(save_dir, '{}.encoder.pt'.format(name)))
    for k, v in decoders.items():
        torch.save(OrderedDict((n, t.cpu()) for n, t in v.state_dict().items()),
                   os.path.join(save_dir, '{}.decoder-{}.pt'.format(name, k)))
    torch.save(optimizer.state_dict(), os.path.join(save_dir, '{}.optimizer.pt'.format(name)))


def load_model(save_dir, name, encoder, decoders, optimizer, device='cuda'):
    """train_vipe_model.py:186-199"""
    import os
    encoder.load_state_dict(torch.load(os.path.join(save_dir, '{}.encoder.pt'.format(name)),
                                       map_location='cpu'))
    for k, decoder in decoders.items():
        decoder.load_state_dict(torch.load(
            os.path.join(save_dir, '{}.decoder-{}.pt'.format(name, k)), map_location='cpu'))
    optimizer.load_state_dict(torch.load(os.path.join(save_dir, '{}.optimizer.pt'.format(name)),
                                         map_location='cpu'))


def get_last_checkpoint(save_dir):
    """train_vipe_model.py:202-209"""
    import os
    import re
    last_epoch = -1
    for fname in os.listdir(save_dir):
        m = re.match(r'epoch(\d+).encoder.pt', fname)
        if m:
            last_epoch = max(int(m.group(1)), last_epoch)
    return last_epoch


def fit(model, train_loaders, val_loaders, save_dir, config, optimizer, num_epochs,
        checkpoint_frequency=25, model_select_contrast=False, model_select_window=1,
        resume=False, weight_3d=1, log=print):
    """The epoch loop of train_vipe_model.main (:312-430) without the video previews: config.json
    (same keys), loss.json (same records), `best_epoch.*` whenever the moving average of the
    validation loss improves, `epochNNNN.*` every `checkpoint_frequency` epochs, resume from the
    last checkpoint. `model` needs `.encoder`, `.decoders` and `.epoch` (Keypoint_EmbeddingModel).
    Returns the loss history."""
    import json
    import os
    import numpy as np
    mv = lambda ls, n, key: float(np.mean([l[key] for l in ls[-n:]]))
    loss_file = os.path.join(save_dir, 'loss.json')
    if resume:
        assert os.path.exists(save_dir)
        last = get_last_checkpoint(save_dir)
        load_model(save_dir, 'epoch{:04d}'.format(last), model.encoder, model.decoders, optimizer)
        start_epoch = last + 1
        with open(loss_file) as fp:
            losses = [x for x in json.load(fp) if x['epoch'] < start_epoch]
        best_val_loss = min([mv(losses[:i], model_select_window, 'val')
                             for i in range(model_select_window, len(losses))] or [float('inf')])
    else:
        missing = [k for k in CONFIG_KEYS if k not in config]
        assert not missing, 'config lacks {}'.format(missing)
        start_epoch = 1
        os.makedirs(save_dir)
        with open(os.path.join(save_dir, 'config.json'), 'w') as fp:
            json.dump({k: config[k] for k in CONFIG_KEYS}, fp, indent=2)
        losses, best_val_loss = [], float('inf')
    for epoch in range(start_epoch, num_epochs + 1):
        tc, tl, dtl = model.epoch(train_loaders, optimizer=optimizer, weight_3d=weight_3d)
        vc, vl, dvl = model.epoch(val_loaders, weight_3d=weight_3d)
        losses.append({'epoch': epoch,
                       'train': tc if model_select_contrast else tl,
                       'val': vc if model_select_contrast else vl,
                       'dataset_train': [('contrast', tc)] + list(dtl.items()),
                       'dataset_val': [('contrast', vc)] + list(dvl.items())})
        mv_avg_val_loss = mv(losses, model_select_window, 'val')
        log('Epoch {} - train loss: {:0.5f}, contra: {:0.3f} [mv-avg: {:0.5f}]'.format(
            epoch, tl, tc, mv(losses, model_select_window, 'train')))
        log('Epoch {} - val loss: {:0.5f}, contra: {:0.3f} [mv-avg: {:0.5f}]'.format(
            epoch, vl, vc, mv_avg_val_loss))
        with open(loss_file, 'w') as fp:
            json.dump(losses, fp, indent=2)
        if mv_avg_val_loss < best_val_loss:
            log('New best epoch!')
            save_model(save_dir, 'best_epoch', model.encoder, model.decoders, optimizer)
        if epoch % checkpoint_frequency == 0:
            log('Saving checkpoint: {}'.format(epoch))
            save_model(save_dir, 'epoch{:04d}'.format(epoch), model.encoder, model.decoders, optimizer)
        best_val_loss = min(mv_avg_val_loss, best_val_loss)
    return losses

"""`ModelTrainer` - drop-in for train_vpd_model.py:53-112, running the native
train step (forward, sum-MSE distillation loss, backward, fused AdamW) and, when
torch.distributed is initialised, summing gradients across ranks with NCCL.

Differences a caller can observe, all deliberate:
  * `get_optimizer` returns `(FusedAdamW, None)`: no GradScaler is needed because
    the kernels keep fp32 master weights / accumulators (the reference only uses
    a scaler for its fp16 autocast path, train_vpd_model.py:100-105);
  * the loss is accumulated on the device and read back once per epoch instead of
    `loss.item()` every step (train_vpd_model.py:93); the returned value is the
    same quantity, sum of losses / number of frames.
"""
import ctypes
import os

import torch

from ._lib import lib, stream_ptr, VpdError

_BUCKET_FN = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64)


class FusedAdamW:
    """torch.optim.AdamW defaults (betas .9/.999, eps 1e-8, weight_decay .01) as one
    kernel launch over the flat parameter arena (K5)."""

    def __init__(self, encoder, lr, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01):
        self.encoder = encoder
        self.lr, self.betas, self.eps, self.weight_decay = lr, betas, eps, weight_decay
        self.step_count = 0
        self._m = self._v = None

    def _state(self):
        p = self.encoder._params
        if self._m is None or self._m.numel() != p.numel():
            self._m = torch.zeros_like(p)
            self._v = torch.zeros_like(p)
        return self._m, self._v

    def step(self):
        enc = self.encoder
        m, v = self._state()
        state, self._bucketed = getattr(self, '_bucketed', None), None
        if state == 'done':
            return          # every bucket of this step went through step_range already
        if state != 'open':
            self.step_count += 1     # ('open': begin_bucketed counted the step, no bucket came)
        with torch.cuda.device(enc._dev):
            net = enc._net
            if net is not None and getattr(net, 'has_grads', False):
                # the bound net's variant also leaves the bf16 operand mirrors of the updated
                # conv weights behind (no weight-packing pass at the start of the next step)
                lib().call('vpd_net_adamw', net.handle, m, v, self.lr, self.betas[0],
                           self.betas[1], self.eps, self.weight_decay, self.step_count, 1.0,
                           stream_ptr(enc._dev))
            else:
                lib().call('vpd_adamw', enc._params, enc._grads, m, v, enc._params.numel(),
                           self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay,
                           self.step_count, 1.0, stream_ptr(enc._dev))
                if net is not None:
                    lib().call('vpd_net_params_changed', net.handle)

    # ---- bucket-wise application (ModelTrainer's gradient-bucket hook) -------------------
    def begin_bucketed(self):
        """A train step is about to run whose gradient buckets will be handed to
        `step_range` as they become final."""
        self._state()
        self.step_count += 1
        self._bucketed = 'open'

    def step_range(self, net, offset, count, finish, stream):
        """AdamW for parameters [offset, offset + count) (one gradient bucket) on `stream`."""
        m, v = self._state()
        lib().call('vpd_net_adamw_range', net.handle, m, v, self.lr, self.betas[0], self.betas[1],
                   self.eps, self.weight_decay, self.step_count, 1.0, offset, count, int(finish),
                   stream.cuda_stream)
        if finish:
            self._bucketed = 'done'

    def zero_grad(self, set_to_none=False):
        pass    # every train step overwrites the whole gradient arena

    def state_dict(self):
        m, v = self._state()
        return {'step': self.step_count, 'exp_avg': m.clone(), 'exp_avg_sq': v.clone(),
                'lr': self.lr, 'betas': self.betas, 'eps': self.eps,
                'weight_decay': self.weight_decay}

    def load_state_dict(self, sd):
        m, v = self._state()
        m.copy_(sd['exp_avg'])
        v.copy_(sd['exp_avg_sq'])
        self.step_count = sd['step']


class FusedSGD:
    """torch.optim.SGD (momentum / dampening / weight_decay / nesterov) as one kernel launch
    over the flat parameter arena."""

    def __init__(self, encoder, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False):
        self.encoder = encoder
        self.lr, self.momentum, self.dampening = lr, momentum, dampening
        self.weight_decay, self.nesterov = weight_decay, nesterov
        self.step_count = 0
        self._buf = None

    def step(self):
        enc = self.encoder
        p = enc._params
        if self.momentum != 0 and (self._buf is None or self._buf.numel() != p.numel()):
            self._buf = torch.zeros_like(p)
            self.step_count = 0
        self.step_count += 1
        with torch.cuda.device(enc._dev):
            lib().call('vpd_sgd', p, enc._grads, self._buf, p.numel(), self.lr, self.momentum,
                       self.dampening, self.weight_decay, int(self.nesterov),
                       int(self.step_count == 1), 1.0, stream_ptr(enc._dev))
            if enc._net is not None:
                lib().call('vpd_net_params_changed', enc._net.handle)

    def zero_grad(self, set_to_none=False):
        pass    # every train step overwrites the whole gradient arena


from . import dp


def _dist():
    return dp.active()


class ModelTrainer:
    """Class for training the encoder. Discarded after training"""

    def __init__(self, encoder, motion):
        self.encoder = encoder.to(encoder.device)
        self.motion = bool(motion)
        if motion:
            encoder._attach_decoder()
        dev = encoder._dev
        self._loss = torch.zeros(1, device=dev, dtype=torch.float64)
        self._copy_stream = torch.cuda.Stream(device=dev)
        # data parallel: gradient buckets are all-reduced on their own stream while the
        # rest of the backward pass is still running (hook: vpd_net_set_bucket_callback)
        self._comm_stream = torch.cuda.Stream(device=dev)
        self._bucket_cb = _BUCKET_FN(self._on_bucket)
        self._hooked = None
        self._pending = []
        self._buckets_seen = []          # (offset, count) handed out by the last train step
        self.overlap_allreduce = os.environ.get('VPD_DP_OVERLAP', '1') != '0'
        # AdamW bucket by bucket on the communication stream while the backward pass of the
        # earlier layers still runs (HBM-bound optimizer under tensor-bound kernels). Opt-in
        # (VPD_BUCKET_ADAMW=1): parity-tested, but measured neutral - the step is SM- and
        # power-bound, so the overlapped optimizer only displaces backward kernels (1 GPU: 3.375
        # vs 3.405 ms at 2.5 % higher clocks; 2 GPUs: 3.502 vs 3.490 ms, e2e 140 k vs 145 k
        # frames/s because of the extra host work per bucket)
        self.bucket_adamw = os.environ.get('VPD_BUCKET_ADAMW', '0') != '0'
        self._step_opt = None
        # data parallel: every replica starts from rank 0's parameters, BN running statistics
        # and counters (identical seeds are NOT assumed - an augmenting loader draws from the
        # same global generators the constructor does, and those must differ per rank)
        dp.broadcast_state([encoder._params, encoder._buffers, encoder._nbt])

    def _on_bucket(self, user, offset, count):
        """Called by the native step when grads[offset:offset+count] are enqueued."""
        dist = _dist() if self.overlap_allreduce else None
        opt = self._step_opt
        if count <= 0 or (dist is None and opt is None):
            return
        self._buckets_seen.append((int(offset), int(count)))
        enc = self.encoder
        cur = torch.cuda.current_stream(enc._dev)
        ev = torch.cuda.Event()
        ev.record(cur)
        self._comm_stream.wait_event(ev)
        with torch.cuda.stream(self._comm_stream):
            if dist is not None:
                work = dp.sum_bucket(enc._grads, offset, count, async_op=True)
                if opt is not None and work is not None:
                    work.wait()          # stream-side: the optimizer below follows the reduce
                else:
                    self._pending.append(work)
            if opt is not None:
                opt.step_range(enc._net, offset, count, finish=(offset == 0),
                               stream=self._comm_stream)

    def _hook(self, net):
        """(Un)install the bucket callback on the bound native net."""
        want = (_dist() is not None and self.overlap_allreduce) or self._step_opt is not None
        key = (net.handle, want)
        if self._hooked != key:
            lib().call('vpd_net_set_bucket_callback', net.handle,
                       ctypes.cast(self._bucket_cb, ctypes.c_void_p) if want else None, None)
            self._hooked = key
        return want

    # ------------------------------------------------------------- one batch
    def _tgt_buffer(self, n, dim):
        """Persistent target buffer: the captured step graphs are keyed by their input / target
        pointers, so batches are copied into fixed buffers instead of handing the loader's
        freshly allocated tensors (a new graph key per batch) to the native step."""
        key = (n, dim)
        bufs = self.__dict__.setdefault('_tgt_bufs', {})
        if key not in bufs:
            if len(bufs) >= 4:
                bufs.clear()
            bufs[key] = torch.empty((n, dim), device=self.encoder._dev, dtype=torch.float32)
        return bufs[key]

    def _run(self, img, tgt, n, train):
        """One batch in the reference loader's format: fp32 NCHW `img`, fp32 `tgt` on the device.
        The layout conversion into the bound net's own input buffer is a launch of its own, so
        the step itself always sees the same pointers (one captured graph per batch size)."""
        enc = self.encoder
        C, H, W = img.shape[-3:]
        if img.dtype != torch.float32 or tgt.dtype != torch.float32:
            raise TypeError('img / emb must be float32')
        with torch.cuda.device(enc._dev):
            if train:
                enc._ensure_grads()
                self._buckets_seen = []
            net = enc._native(H, W, n)
            self._overlapped = self._hook(net) if train else False
            stem = lib().call('vpd_net_stem_input', net.handle)
            st = stream_ptr(enc._dev)
            lib().call('vpd_nchw_to_stem', img.contiguous(), stem, n, C, H, W, st)
            buf = self._tgt_buffer(n, tgt.shape[1])
            buf.copy_(tgt, non_blocking=True)
            if train:
                lib().call('vpd_net_train_step', net.handle, None, stem, buf, n, self._loss, st)
            else:
                lib().call('vpd_net_eval_loss', net.handle, None, stem, buf, n, self._loss, None, st)

    def train_step_stem(self, stem, tgt, n, height, width, optimizer):
        """Device-resident fast path: input already in the network layout (K1
        `assemble_stem`), target fp32 [n, T] on the device."""
        enc = self.encoder
        with torch.cuda.device(enc._dev):
            enc._ensure_grads()
            self._buckets_seen = []
            net = enc._native(height, width, n)
            self._arm_optimizer(optimizer, True)
            self._overlapped = self._hook(net)
            lib().call('vpd_net_train_step', net.handle, None, stem, tgt, n, self._loss,
                       stream_ptr(enc._dev))
            self._sync_grads()
            optimizer.step()
            self._step_opt = None

    def dp_self_check(self, img, tgt, n, rtol=1e-4):
        """Data-parallel correctness of ONE step, judged bucket by bucket: the
        gradient arena the trainer's path leaves behind (bucketed all-reduce overlapped with the
        backward pass, or the single call) must equal the SUM over ranks of the gradients every
        rank computes on its own for the same batch. Runs the step twice on `img`/`tgt` (fp32
        NCHW batch + targets on the device): once without any exchange, then through the
        trainer; the two runs agree bit for bit in everything but the weight-gradient kernels'
        fp32 atomics (integer BatchNorm statistics), so the bar is fp32 rounding, and a bucket
        that was never reduced misses it by four orders of magnitude. Does not touch the
        parameters or the BN running statistics. -> dp.check_bucket_sums(...) + {'buckets', 'overlapped'}; every rank must
        call it (collectives inside)."""
        dist = _dist()
        enc = self.encoder
        H, W = img.shape[-2:]
        self._step_opt = None                # gradients only: the optimizer must not run here
        loss_keep = self._loss.clone()
        with torch.cuda.device(enc._dev):
            enc._ensure_grads()
            net = enc._native(H, W, n)
            keep = (enc._buffers.clone(), enc._nbt.clone())     # BN running state is restored
            lib().call('vpd_net_set_bucket_callback', net.handle, None, None)
            self._hooked = None
            lib().call('vpd_net_train_step', net.handle, img, None, tgt, n, self._loss,
                       stream_ptr(enc._dev))
            expect = enc._grads.clone()
            if dist is not None:
                dist.all_reduce(expect, op=dist.ReduceOp.SUM)
            self._run(img, tgt, n, True)
            self._sync_grads()
            torch.cuda.synchronize(enc._dev)
            buckets = list(self._buckets_seen) if self._overlapped else []
            if not buckets:
                buckets = [(0, enc._grads.numel())]
            res = dp.check_bucket_sums(enc._grads, expect, buckets, rtol=rtol)
            enc._buffers.copy_(keep[0])
            enc._nbt.copy_(keep[1])
        self._loss.copy_(loss_keep)
        res['buckets'] = len(buckets)
        res['overlapped'] = bool(self._overlapped)
        return res

    def stem_buffer(self, n, height, width):
        """Raw device pointer of the bound net's own input buffer (so assembly can
        write into it without an extra copy)."""
        enc = self.encoder
        with torch.cuda.device(enc._dev):
            enc._ensure_grads()
            net = enc._native(height, width, n)
            return lib().call('vpd_net_stem_input', net.handle)

    def _sync_grads(self):
        dist = _dist()
        if getattr(self, '_overlapped', False):
            # buckets (all-reduce and / or optimizer ranges) were launched from the callback on
            # the communication stream: make the compute stream wait for them
            for work in self._pending:
                if work is not None:
                    work.wait()
            self._pending = []
            torch.cuda.current_stream(self.encoder._dev).wait_stream(self._comm_stream)
        elif dist is not None:
            dp.sum_gradients(self.encoder._grads)

    def _arm_optimizer(self, optimizer, train):
        """Bucket-wise AdamW for this step? Only with the package's own optimizer; a data-
        parallel run without the overlapped exchange reduces the whole arena after the step,
        so the optimizer has to wait for that."""
        ok = (train and self.bucket_adamw and isinstance(optimizer, FusedAdamW)
              and (_dist() is None or self.overlap_allreduce))
        self._step_opt = optimizer if ok else None
        if ok:
            optimizer.begin_bucketed()

    def _run_u8(self, raw, tgt, train):
        """One batch of raw uint8 crops (see `_stage`): K1 assembly straight into the bound
        net's input buffer, then the native step."""
        from .assemble import assemble_stem
        enc = self.encoder
        rgb, flow, flip, teacher = raw['rgb_u8'], raw.get('flow_u8'), raw.get('flip'), raw.get('teacher')
        index = raw.get('index')          # frames = rows `index` of device-resident pools
        n, H, W, _ = rgb.shape
        if index is not None:
            if index.device != rgb.device or index.dtype != torch.int32:
                raise ValueError("'index' must be an int32 tensor on the pools' device")
            n = int(index.shape[0])
        if (flow is not None) != bool(enc.use_flow):
            raise AssertionError('Wrong number of channels for RGB' + (' + flow' if enc.use_flow else ''))
        ms = raw.get('rgb_mean_std', getattr(self, 'rgb_mean_std', None))
        if ms is None:
            raise ValueError("uint8 batches need 'rgb_mean_std' (in the batch or on the trainer)")
        expect = 2 * enc.emb_dim if self.motion else enc.emb_dim
        with torch.cuda.device(enc._dev):
            if train:
                enc._ensure_grads()
                self._buckets_seen = []
            net = enc._native(H, W, n)
            stem = lib().call('vpd_net_stem_input', net.handle)
            if teacher is not None:
                if teacher.shape[-1] != expect:
                    raise ValueError('target dim {} != {}'.format(teacher.shape[-1], expect))
                tgt = self._tgt_buffer(n, expect)
                assemble_stem(stem, rgb, flow, ms, flip=flip, teacher=teacher, tgt=tgt, index=index)
            else:
                if tgt.shape[1] != expect:
                    raise ValueError('target dim {} != {}'.format(tgt.shape[1], expect))
                buf = self._tgt_buffer(n, expect)
                buf.copy_(tgt, non_blocking=True)
                tgt = buf
                assemble_stem(stem, rgb, flow, ms, flip=flip, index=index)
            self._overlapped = self._hook(net) if train else False
            if train:
                lib().call('vpd_net_train_step', net.handle, None, stem, tgt, n, self._loss,
                           stream_ptr(enc._dev))
            else:
                lib().call('vpd_net_eval_loss', net.handle, None, stem, tgt, n, self._loss, None,
                           stream_ptr(enc._dev))
        return n

    def _stage_u8(self, batch, slot):
        """Raw batch {'rgb_u8': uint8 [B,H,W,3], 'flow_u8': uint8 [B,H,W,>=2] (flow models),
        'flip': uint8 [B] (optional), and either 'emb': fp32 [B,T] (teacher row already chosen)
        or 'teacher': fp32 [B,2,T] (row = flip bit, chosen on the device),
        'rgb_mean_std': ((m,m,m),(s,s,s)) (or set `trainer.rgb_mean_std`)}: 4x fewer bytes over
        PCIe than the fp32 batch; normalisation / stacking / flip run in the K1 kernel."""
        dev = self.encoder._dev
        names = [k for k in ('rgb_u8', 'flow_u8', 'flip', 'emb', 'teacher', 'index') if batch.get(k) is not None]
        extra = {k: batch[k] for k in ('rgb_mean_std',) if k in batch}
        if batch.get('index') is not None and batch['rgb_u8'].device != dev:
            raise ValueError("'index' batches gather from device-resident pools")
        if all(batch[k].device == dev for k in names):     # already resident: no staging ring
            raw = {k: batch[k] for k in names}
            raw.update(extra)
            return raw, raw.get('emb'), None, None
        # two device slots sized for the largest batch seen; shorter batches use slices
        n = batch['rgb_u8'].shape[0]
        key = ('u8',) + tuple((k, tuple(batch[k].shape[1:]), batch[k].dtype) for k in names)
        ring = getattr(self, '_ring', None)
        if ring is None or ring.get('key') != key or ring['cap'] < n:
            cap = n if ring is None or ring.get('key') != key else max(ring['cap'], n)
            ring = {'key': key, 'cap': cap, 'free': [None, None],
                    'buf': [{k: torch.empty((cap,) + tuple(batch[k].shape[1:]), device=dev,
                                            dtype=batch[k].dtype) for k in names}
                            for _ in range(2)]}
            self._ring = ring
        with torch.cuda.stream(self._copy_stream):
            if ring['free'][slot] is not None:
                self._copy_stream.wait_event(ring['free'][slot])
            for k in names:
                ring['buf'][slot][k][:n].copy_(batch[k], non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        raw = {k: ring['buf'][slot][k][:n] for k in names}
        raw.update(extra)
        return raw, raw.get('emb'), ev, slot

    def _stage(self, batch, slot):
        """Start the host->device copies of a batch on the copy stream, into one of
        two persistent device staging buffers (no per-step allocation)."""
        if batch is None:
            return None
        if 'rgb_u8' in batch:
            return self._stage_u8(batch, slot)
        dev = self.encoder._dev
        img_h, emb_h = batch['img'], batch['emb']
        if img_h.device == dev and emb_h.device == dev and img_h.dtype == torch.float32:
            return img_h, emb_h, None, None            # already resident: no staging ring
        # ring sized for the largest batch seen (a shorter last batch of an epoch uses a slice)
        ring = getattr(self, '_ring', None)
        need = (tuple(img_h.shape[1:]), tuple(emb_h.shape[1:]))
        if (ring is None or ring.get('key') != need or ring['cap'] < img_h.shape[0]):
            cap = img_h.shape[0] if ring is None or ring.get('key') != need else \
                max(ring['cap'], img_h.shape[0])
            ring = {'key': need, 'cap': cap, 'free': [None, None],
                    'img': [torch.empty((cap,) + need[0], device=dev, dtype=torch.float32)
                            for _ in range(2)],
                    'emb': [torch.empty((cap,) + need[1], device=dev, dtype=torch.float32)
                            for _ in range(2)]}
            self._ring = ring
        n = img_h.shape[0]
        with torch.cuda.stream(self._copy_stream):
            if ring['free'][slot] is not None:         # previous consumer of this slot done?
                self._copy_stream.wait_event(ring['free'][slot])
            ring['img'][slot][:n].copy_(img_h, non_blocking=True)
            ring['emb'][slot][:n].copy_(emb_h, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self._copy_stream)
        return ring['img'][slot][:n], ring['emb'][slot][:n], ev, slot

    def epoch(self, data_loader, optimizer=None, scaler=None, progress_cb=None):
        enc = self.encoder
        train = optimizer is not None
        enc.eval() if optimizer is None else enc.train()
        self._loss.zero_()
        epoch_n = 0
        it = iter(data_loader)
        step_i = 0
        nxt = self._stage(next(it, None), 0)
        cur_stream = torch.cuda.current_stream(enc._dev)
        while nxt is not None:
            img, emb, ev, slot = nxt
            step_i += 1
            nxt = self._stage(next(it, None), step_i % 2)   # overlap next copy with this step
            if ev is not None:
                cur_stream.wait_event(ev)
            self._arm_optimizer(optimizer, train)
            if isinstance(img, dict):                       # raw uint8 crops: K1 on the device
                n = self._run_u8(img, emb, train)
            else:
                img = img.contiguous()
                emb = emb.contiguous()
                n = img.shape[0]
                enc._check_channels(img)
                expect = 2 * enc.emb_dim if self.motion else enc.emb_dim
                if emb.shape[1] != expect:
                    raise ValueError('target dim {} != {}'.format(emb.shape[1], expect))
                self._run(img, emb, n, train)
            if slot is not None:
                done = torch.cuda.Event()
                done.record(cur_stream)
                self._ring['free'][slot] = done
            if train:
                self._sync_grads()
                optimizer.step()
                optimizer.zero_grad()
                self._step_opt = None
            epoch_n += n
            if progress_cb is not None:
                progress_cb(n)
        if epoch_n == 0:
            return float('nan')
        # data parallel: the loss over the GLOBAL batch, the same number on every rank (so every
        # rank takes the same model-selection decisions in fit())
        return dp.epoch_loss(self._loss, epoch_n)

    def get_optimizer(self, learning_rate):
        return FusedAdamW(self.encoder, learning_rate), None

    def save_model(self, save_dir, name):
        sd = {k: v.cpu() for k, v in self.encoder.state_dict().items()}
        torch.save(sd, os.path.join(save_dir, '{}.encoder.pt'.format(name)))
        if self.motion:
            dsd = {k: v.cpu() for k, v in self.encoder.decoder_state_dict().items()}
            torch.save(dsd, os.path.join(save_dir, '{}.decoder.pt'.format(name)))

"""The keypoint (VIPE*) teacher's APPLY path on the B200: `FCResNet` encoder and
`Keypoint_EmbeddingModel.embed`, the producer of the `<video>.emb.pkl` teacher files that the
student's target construction (`vpd_b200/targets.py`, reference single_frame.py:208-273) reads.

Reference: models/module.py:159-204 (FcResidualBlock: Linear, BatchNorm1d, ReLU, Dropout,
Linear, BatchNorm1d, ReLU, Dropout, then `x2 - x`; FCResNet: Linear + ReLU, the blocks, a last
Linear), models/keypoint.py:14-35,128-160 (`_BaseModel`, `_predict`, `embed`),
apply_vipe_model.py:37-69,133-204 (`mean_embs_by_frame`, `load_embedding_model`, the per-video
loop). Same constructor arguments, same `state_dict()` keys / shapes / dtypes (checkpoints are
interchangeable), same initialisation draw for draw (`vpd_b200/init.py::fcresnet_state`).

Eval forward on the device, all through the C ABI (no torch arithmetic, no fallback):
    poses fp32 [n, in_dim] --vpd_rows_to_bf16--> bf16 [n, 64 (128 with bones)]
    Linear(in, hidden) + ReLU                vpd_conv2d_fwd (1x1 over [n,1,1,64]; bias = shift)
    per block: Linear + BN1d(eval) + ReLU    vpd_conv2d_fwd, bias and BN folded by vpd_bn_fold
               Linear + BN1d(eval) + ReLU    vpd_conv2d_fwd
               x2 - x                        vpd_axpby_bf16
    Linear(hidden, emb_dim)                  vpd_linear_rows_f32 (fp32 weights, fp32 output)
Dropout is the identity in eval mode. Hidden activations are bf16 (tensor-core operands);
the embedding is accumulated and returned in fp32.

Training the teacher (`Keypoint_EmbeddingModel.epoch`: three weight-sharing encoder passes,
hinge + MSE losses, models/keypoint.py:38-126) and the FC 3-D pose decoder live in
`vpd_b200/keypoint_train.py`; `FCResNet.forward` itself stays eval-only (the training forward
needs the saved activations the trainer keeps).
"""
import json
import os
import pickle
from collections import OrderedDict, defaultdict

import numpy as np
import torch

from . import init as _init
from ._lib import lib, stream_ptr, VpdError

NUM_COCO_KEYPOINTS = 13      # vipe_dataset/dataset_base.py (COCO keypoints used by VIPE)
NUM_COCO_BONES = 12
EMBED_BATCH_SIZE = 250       # apply_vipe_model.py:19 (the reference's chunk; we take any n)
_MAX_ROWS = 1 << 16          # rows per launch group (bounds the activation buffers)
_BN_EPS = 1e-5


class FCResNet:
    """models/module.py:192-204, eval forward on the GPU."""

    def __init__(self, in_dim, out_dim, num_blocks, hidden_dim, dropout=0.3):
        if out_dim is None:
            raise NotImplementedError('FCResNet without the output Linear (the decoder trunk) '
                                      'is not part of the CUDA path yet')
        if hidden_dim % 64 != 0 or in_dim > 512 or out_dim > 64:
            raise NotImplementedError(
                'CUDA path: hidden_dim % 64 == 0, in_dim <= 512, out_dim <= 64 '
                '(got in={}, hidden={}, out={})'.format(in_dim, hidden_dim, out_dim))
        self.in_dim, self.out_dim = in_dim, out_dim
        self._cin = (in_dim + 63) // 64 * 64          # pose columns padded to whole channel chunks
        self.num_blocks, self.hidden_dim, self.dropout = num_blocks, hidden_dim, dropout
        self.training = True                          # nn.Module default
        self._sd = _init.fcresnet_state(in_dim, out_dim, num_blocks, hidden_dim)
        self._dev = None
        self._prepared = None

    # ---- nn.Module surface -------------------------------------------------------------
    def to(self, device):
        dev = torch.device('cuda' if str(device) == 'cuda' else device)
        if dev.type != 'cuda':
            raise VpdError("vpd_b200 needs a CUDA device; got '{}' (no CPU fallback)".format(device))
        if dev.index is None:
            dev = torch.device('cuda', torch.cuda.current_device())
        lib()
        self._dev = dev
        self._sd = OrderedDict((k, v.to(dev)) for k, v in self._sd.items())
        self._prepared = None
        return self

    def train(self, mode=True):
        self.training = bool(mode)
        return self

    def eval(self):
        return self.train(False)

    def state_dict(self):
        return OrderedDict((k, v.detach().clone()) for k, v in self._sd.items())

    def load_state_dict(self, sd, strict=True):
        missing = [k for k in self._sd if k not in sd]
        unexpected = [k for k in sd if k not in self._sd]
        if strict and (missing or unexpected):
            raise RuntimeError('Error(s) in loading state_dict for FCResNet: missing {}, '
                               'unexpected {}'.format(missing, unexpected))
        for k, v in sd.items():
            if k in self._sd:
                if tuple(v.shape) != tuple(self._sd[k].shape):
                    raise RuntimeError('size mismatch for {}: {} vs {}'.format(
                        k, tuple(v.shape), tuple(self._sd[k].shape)))
                # in place: during training the entries are views into the flat arena
                self._sd[k].copy_(v.detach().to(device=self._sd[k].device, dtype=self._sd[k].dtype))
        self._prepared = None
        self._version = getattr(self, '_version', 0) + 1    # KeypointTrainCore re-packs its mirrors

    def parameters(self):
        return [v for k, v in self._sd.items()
                if not k.endswith(('running_mean', 'running_var', 'num_batches_tracked'))]

    # ---- device-side constants derived from the parameters -------------------------------
    def _prepare(self):
        if self._prepared is not None:
            return self._prepared
        if self._dev is None:
            raise VpdError('FCResNet: call .to(cuda device) first (no CPU fallback)')
        L, dev, H = lib(), self._dev, self.hidden_dim
        st = stream_ptr(dev)
        bf = lambda *s: torch.empty(s, device=dev, dtype=torch.bfloat16)
        f32 = lambda *s: torch.empty(s, device=dev, dtype=torch.float32)
        sd = self._sd
        layers = []
        # Linear(in, hidden): input columns zero-padded to a multiple of 64
        cin = self._cin
        w0 = torch.zeros((H, cin), device=dev, dtype=torch.float32)
        w0[:, :self.in_dim] = sd['layers.0.weight']
        wt = bf(H * cin)
        L.call('vpd_pack_conv_weight', w0, wt, None, H, cin, 1, st)
        layers.append((wt, torch.ones(H, device=dev), sd['layers.0.bias'].contiguous(), cin))
        for i in range(self.num_blocks):
            p = 'layers.{}.block'.format(2 + i)
            for lin, bn in ((0, 1), (4, 5)):
                wt = bf(H * H)
                L.call('vpd_pack_conv_weight', sd['{}.{}.weight'.format(p, lin)].contiguous(), wt,
                       None, H, H, 1, st)
                scale, shift = f32(H), f32(H)
                b = '{}.{}'.format(p, bn)
                L.call('vpd_bn_fold', sd[b + '.weight'], sd[b + '.bias'], sd[b + '.running_mean'],
                       sd[b + '.running_var'], sd['{}.{}.bias'.format(p, lin)], _BN_EPS, scale,
                       shift, H, st)
                layers.append((wt, scale, shift, H))
        last = 'layers.{}'.format(2 + self.num_blocks)
        self._prepared = (layers, sd[last + '.weight'].contiguous(), sd[last + '.bias'].contiguous())
        return self._prepared

    def forward(self, x):
        """x fp32 [n, in_dim] (device or host) -> fp32 [n, out_dim] on the device. Eval mode
        only: the training forward (batch statistics, dropout) belongs to the teacher's
        training path, which is not built."""
        if self.training:
            raise NotImplementedError('FCResNet.forward in training mode (teacher training) is '
                                      'not part of the CUDA path yet; call .eval()')
        layers, w_out, b_out = self._prepare()
        L, dev, H = lib(), self._dev, self.hidden_dim
        x = x.to(device=dev, dtype=torch.float32).contiguous()
        if x.dim() != 2 or x.shape[1] != self.in_dim:
            raise ValueError('expected [n, {}] poses, got {}'.format(self.in_dim, tuple(x.shape)))
        n = x.shape[0]
        out = torch.empty((n, self.out_dim), device=dev, dtype=torch.float32)
        st = stream_ptr(dev)
        m = min(n, _MAX_ROWS)
        xb = torch.empty((m, self._cin), device=dev, dtype=torch.bfloat16)
        h, z1, z2 = (torch.empty((m, H), device=dev, dtype=torch.bfloat16) for _ in range(3))

        def linear(src, dst, rows, layer):
            wt, scale, shift, cin = layer
            L.call('vpd_conv2d_fwd', src, wt, dst, rows, 1, 1, cin, H, 1, 1, 0, scale, shift,
                   None, 1, None, st)

        for r0 in range(0, n, _MAX_ROWS):
            rows = min(_MAX_ROWS, n - r0)
            L.call('vpd_rows_to_bf16', x[r0:r0 + rows], xb, rows, self.in_dim, self._cin, st)
            linear(xb, h, rows, layers[0])
            for i in range(self.num_blocks):
                linear(h, z1, rows, layers[1 + 2 * i])
                linear(z1, z2, rows, layers[2 + 2 * i])
                L.call('vpd_axpby_bf16', z2, 1.0, h, -1.0, z1, rows * H, st)     # x2 - x
                h, z1 = z1, h
            L.call('vpd_linear_rows_f32', h, w_out, b_out, out[r0:r0 + rows], rows, H,
                   self.out_dim, st)
        return out

    __call__ = forward


def batch_mulitplexer(data_loaders):
    """models/util.py:5-23 (evaluation order): yields (name, batch) until every loader is drained;
    the next loader is drawn with `random.choices` weighted by the batches it has left, so the
    draws consumed from Python's generator equal the reference's."""
    import random
    live = [[name, len(loader), iter(loader)] for name, loader in data_loaders]
    while live:
        pick = random.choices(list(range(len(live))), k=1, weights=[e[1] for e in live])[0]
        entry = live[pick]
        try:
            batch = next(entry[2])
        except StopIteration:
            raise Exception('loader {} ended before its announced length'.format(entry[0]))
        yield entry[0], batch
        entry[1] -= 1
        if entry[1] == 0:
            del live[pick]


def batch_zipper(data_loaders):
    """models/util.py:26-47 (training order): round i yields one batch from every loader, except
    that a loader with fewer batches than the longest sits out `deficit` rounds drawn without
    replacement from numpy's global generator (same call, same order as the reference)."""
    rounds = max(len(loader) for _, loader in data_loaders)
    sits_out = {}
    for name, loader in data_loaders:
        deficit = rounds - len(loader)
        if deficit > 0:
            sits_out[name] = set(np.random.choice(np.arange(rounds), deficit, replace=False).tolist())
    iters = [(name, iter(loader)) for name, loader in data_loaders]
    for i in range(rounds):
        yield [(name, next(it)) for name, it in iters if i not in sits_out.get(name, ())]


class Keypoint_EmbeddingModel:
    """models/keypoint.py:14-35,38-160. `decoders` = {} or {'3d': FCPoseDecoder}
    (vpd_b200.keypoint_train.FCPoseDecoder; the FCResNet decoder variant is not built)."""

    def __init__(self, encoder, decoders, device):
        self.encoder = encoder
        self.decoders = decoders
        self.device = device
        self.encoder.to(device)
        if decoders:
            from .keypoint_train import FCPoseDecoder
            if set(decoders) != {'3d'} or not isinstance(decoders['3d'], FCPoseDecoder):
                raise NotImplementedError("decoders must be {} or {'3d': FCPoseDecoder}")
        self._core_obj = None

    def _core(self):
        if self._core_obj is None:
            from .keypoint_train import KeypointTrainCore
            self._core_obj = KeypointTrainCore(self.encoder, self.decoders.get('3d'),
                                               self.encoder._dev)
        return self._core_obj

    def get_optimizer(self, learning_rate):
        """AdamW over encoder + decoder parameters (train_vipe_model.py:164-169,312-314) as one
        fused launch over the flat arena. (The reference script builds torch.optim.AdamW over
        `get_model_params(encoder, decoders)` itself; build this one instead.)"""
        from .keypoint_train import KeypointAdamW
        return KeypointAdamW(self._core(), learning_rate)

    def _to_dev(self, batch):
        dev = self.encoder._dev
        out = {}
        n = batch['pose1'].shape[0]
        for k in ('pose1', 'pose2', 'pose_neg'):
            if k in batch:
                out[k] = torch.as_tensor(batch[k]).to(dev, torch.float32).reshape(n, -1).contiguous()
        if 'pose_neg_is_valid' in batch:
            out['pose_neg_is_valid'] = torch.as_tensor(batch['pose_neg_is_valid']).to(
                dev, torch.float32).reshape(n).contiguous()
        if 'kp_features' in batch:
            out['kp_features'] = torch.as_tensor(batch['kp_features']).float().to(dev).reshape(
                n, -1).contiguous()
        return out

    def epoch(self, data_loaders, optimizer=None, scaler=None, progress_cb=None, weight_3d=1,
              dropout_masks=None):
        """models/keypoint.py:38-126. data_loaders: [(dataset_name, loader of batch dicts)].
        Returns (contrastive loss / n, loss / n, {dataset: loss / n}). `scaler` is ignored (fp32
        master weights, bf16 range). dropout_masks: optional iterator yielding, per dataset
        batch, the keep masks [pass][2 * num_blocks] uint8 [n, hidden] (tests)."""
        from collections import Counter
        train = optimizer is not None
        self.encoder.train(train)
        core = self._core()
        core.weights_dirty = True
        self.encoder._prepared = None
        dataset_losses, dataset_contra_losses, dataset_counts = Counter(), Counter(), Counter()
        batches = batch_zipper(data_loaders) if train else (
            (x,) for x in batch_mulitplexer(data_loaders))
        for zipped_batch in batches:
            zipped_batch = [(name, self._to_dev(b)) for name, b in zipped_batch]
            batch_n = sum(b['pose1'].shape[0] for _, b in zipped_batch)
            for dataset_name, batch in zipped_batch:
                core.loss_sums.zero_()
                masks = next(dropout_masks) if dropout_masks is not None else None
                n = core.dataset_step(batch, dataset_name, 1.0 / batch_n, weight_3d, masks=masks,
                                      train=train)
                contra_loss, loss = core.loss_sums.tolist()
                if contra_loss > 0:
                    dataset_contra_losses[dataset_name] += contra_loss
                dataset_losses[dataset_name] += loss
                dataset_counts[dataset_name] += n
            if train:
                optimizer.step()
                optimizer.zero_grad()
                self.encoder._prepared = None
            if progress_cb is not None:
                progress_cb(batch_n)
        epoch_n = sum(dataset_counts.values())
        return (sum(dataset_contra_losses.values()) / epoch_n,
                sum(dataset_losses.values()) / epoch_n,
                {k: v / dataset_counts[k] for k, v in dataset_losses.items()})

    def _predict(self, pose, get_emb, decoder_target=None):
        assert get_emb or decoder_target is not None, 'Nothing to predict'
        if not isinstance(pose, torch.Tensor):
            pose = torch.FloatTensor(np.asarray(pose))
        if len(pose.shape) == 2:
            pose = pose.unsqueeze(0)
        self.encoder.eval()
        n = pose.shape[0]
        emb = self.encoder(pose.reshape(n, -1))
        if decoder_target is None:
            return emb.cpu().numpy(), None
        core = self._core()
        core._refresh_mirrors()
        pred, _ = core.decoder_forward(emb, decoder_target)
        tdim = dict(self.decoders['3d'].target_dims)[decoder_target]
        pred = pred[:, :tdim].float().cpu().numpy()
        return (emb.cpu().numpy() if get_emb else None), pred

    def embed(self, pose):
        return self._predict(pose, get_emb=True)[0]

    def predict3d(self, pose, decoder_target):
        return self._predict(pose, get_emb=False, decoder_target=decoder_target)[1]

    def embed_and_predict3d(self, pose, decoder_target):
        return self._predict(pose, get_emb=True, decoder_target=decoder_target)


def load_embedding_model(model_dir, model_epoch=None, device='cuda'):
    """apply_vipe_model.py:133-162: config.json + `<name>.encoder.pt` -> (model, embed_bones)"""
    with open(os.path.join(model_dir, 'config.json')) as fp:
        params = json.load(fp)
    embed_bones = params['embed_bones']
    name = 'best_epoch' if model_epoch is None else 'epoch{:04d}'.format(model_epoch)
    encoder = FCResNet((NUM_COCO_KEYPOINTS + NUM_COCO_BONES if embed_bones
                        else NUM_COCO_KEYPOINTS) * 3,
                       params['embedding_dim'], *params['encoder_arch'])
    encoder.load_state_dict(torch.load(os.path.join(model_dir, name + '.encoder.pt'),
                                       map_location='cpu'))
    return Keypoint_EmbeddingModel(encoder, {}, device), embed_bones


def mean_embs_by_frame(pred_embs, flip):
    """apply_vipe_model.py:37-69: one entry per frame, sorted by frame number. Several
    detections in a frame are averaged (meta = lowest score, is_mean); with `flip` the entry is
    the [2, D] stack (unflipped, flipped) and carries the unflipped side's meta."""
    if not pred_embs:
        return []
    shape = pred_embs[-1][1].shape
    per_frame = {}
    for frame_num, emb, meta in pred_embs:
        per_frame.setdefault(frame_num, []).append((emb, meta))

    def reduce(items):
        if len(items) == 1:
            e, meta = items[0]
        else:
            e = np.mean([it[0] for it in items], axis=0)
            meta = {'kp_score': min(it[1]['kp_score'] for it in items), 'is_mean': True}
        assert e.shape == shape
        return e, meta

    out = []
    for frame_num, items in per_frame.items():
        if flip:
            e, meta = reduce([it for it in items if not it[1]['is_flip']])
            e_flip, _ = reduce([it for it in items if it[1]['is_flip']])
            out.append((frame_num, np.stack((e, e_flip)), meta))
        else:
            e, meta = reduce(items)
            out.append((frame_num, e, meta))
    out.sort(key=lambda x: x[0])
    return out


def embed_video(model, frames, scores, is_flip, poses, flip=True, allow_many_per_frame=False):
    """The body of apply_vipe_model.main's loop (:181-201) for one video's arrays (what its
    `VideoDataset` yields): -> the list stored in `<video>.emb.pkl`. The whole video is
    embedded in one call instead of chunks of 250."""
    frames, scores, is_flip = np.asarray(frames), np.asarray(scores), np.asarray(is_flip)
    if len(frames) == 0:
        return []
    batch_embs = model.embed(poses)
    embs = [(frames[j].item(), batch_embs[j, :],
             {'kp_score': scores[j].item(), 'is_mean': False, 'is_flip': is_flip[j].item()})
            for j in range(batch_embs.shape[0])]
    if not allow_many_per_frame:
        embs = mean_embs_by_frame(embs, flip)
    return embs


def write_embs(out_dir, video_name, embs):
    """apply_vipe_model.py:171-178 + util/io.py:35-37 (store_pickle)"""
    if embs and video_name is not None and out_dir is not None:
        os.makedirs(out_dir, exist_ok=True)
        with open(os.path.join(out_dir, '{}.emb.pkl'.format(video_name)), 'wb') as fp:
            pickle.dump(embs, fp)

"""`RGBF_EmbeddingModel` - drop-in for the reference's models/rgb.py:46-86, backed by
the sm_100a kernels in libvpd_b200.so.

Same constructor `(model_arch, emb_dim, use_flow, device, pretrained=False)`, same
attributes (`device`, `use_flow`, `emb_dim`), `embed()` / `forward()` /
`state_dict()` / `load_state_dict()` / `to()` / `train()` / `eval()` /
`parameters()`, and `state_dict()` has the reference's 218 `resnet.*` keys with the
reference's shapes and dtypes, so checkpoints are interchangeable
(`train_vpd_model.py:107-112`, `apply_vpd_model.py:141-144`).

Device memory is owned here as torch tensors (flat fp32 arenas + one scratch
workspace); all arithmetic happens in the CUDA library through the C ABI. There is
no PyTorch or CPU fallback: constructing the model without a CUDA device raises.
"""
import ctypes
from collections import OrderedDict

import numpy as np
import torch

from . import init as _init
from ._lib import lib, stream_ptr, VpdError

_MAX_CHUNK = 512        # frames per forward launch group in embed()


class _NativeNet:
    """A vpd_net handle + its workspace for one (H, W, max_batch, motion)."""

    def __init__(self, arch, emb_dim, in_channels, height, width, max_batch, motion):
        self.key = (height, width, max_batch, bool(motion))
        self.handle = lib().call('vpd_net_create', arch.encode(), emb_dim, in_channels, height,
                                 width, max_batch, int(bool(motion)))
        if not self.handle:
            raise VpdError('vpd_net_create failed: ' + lib().last_error())
        self.max_batch = max_batch
        self.workspace = None

    def tensor_table(self):
        L = lib()
        out = []
        name = ctypes.create_string_buffer(128)
        arena, layout, ndim = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        offset = ctypes.c_int64()
        shape = (ctypes.c_int64 * 4)()
        for i in range(L.call('vpd_net_num_tensors', self.handle)):
            L.call('vpd_net_tensor_info', self.handle, i, name, 128, ctypes.byref(arena),
                   ctypes.byref(offset), ctypes.byref(layout), ctypes.byref(ndim), shape)
            out.append((name.value.decode(), arena.value, offset.value, layout.value,
                        tuple(shape[k] for k in range(ndim.value))))
        return out

    def close(self):
        if self.handle:
            lib().call('vpd_net_destroy', self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class RGBF_EmbeddingModel:
    """Basic embedding model with single frame features (B200-native)."""

    def __init__(self, model_arch, emb_dim, use_flow, device, pretrained=False):
        if pretrained:
            raise NotImplementedError(
                'pretrained=True needs the torchvision ImageNet checkpoint (no network here); '
                'load one with load_state_dict() instead')
        if 'effnet' in model_arch:
            raise NotImplementedError('efficientnet encoders are not part of the CUDA path')
        _init.blocks(model_arch)            # validates the architecture name
        if not torch.cuda.is_available():
            raise VpdError('vpd_b200 needs a CUDA device (sm_100a); there is no CPU fallback')
        lib()                               # fail loudly if the extension is missing
        self.model_arch = model_arch
        self.emb_dim = emb_dim
        self.use_flow = use_flow
        self.device = device
        self._dev = torch.device('cuda' if str(device) == 'cuda' else device)
        if self._dev.type != 'cuda':
            raise VpdError("device must be a CUDA device, got '{}'".format(device))
        if self._dev.index is None:
            self._dev = torch.device('cuda', torch.cuda.current_device())
        self.training = True                # nn.Module default
        self._in_channels = 5 if use_flow else 3
        self._motion = False
        self._net = None
        self._layout_net = _NativeNet(model_arch, emb_dim, self._in_channels, 128, 128, 1, False)
        self._alloc_arenas()
        self._write_state(_init.encoder_state(model_arch, emb_dim, use_flow))

    # ------------------------------------------------------------------ arenas
    def _alloc_arenas(self, keep=None):
        L = lib()
        h = self._layout_net.handle
        self._table = self._layout_net.tensor_table()
        self._n_params = L.call('vpd_net_param_count', h)
        self._n_conv = L.call('vpd_net_conv_param_count', h)
        n_buf = L.call('vpd_net_buffer_count', h)
        n_bn = L.call('vpd_net_num_bn', h)
        with torch.cuda.device(self._dev):
            params = torch.zeros(self._n_params, device=self._dev, dtype=torch.float32)
            if keep is not None:            # growing the arena (decoder attached later)
                params[:keep.numel()] = keep
                self._buffers = self._buffers
                self._nbt = self._nbt
            else:
                self._buffers = torch.zeros(n_buf, device=self._dev, dtype=torch.float32)
                self._nbt = torch.zeros(n_bn, device=self._dev, dtype=torch.int64)
            self._params = params
            self._grads = None

    def _ensure_grads(self):
        if self._grads is None or self._grads.numel() != self._n_params:
            self._grads = torch.zeros(self._n_params, device=self._dev, dtype=torch.float32)
            self._net = None                # rebind with the gradient arena
        return self._grads

    def _attach_decoder(self):
        """ModelTrainer(motion=True): the FCNet decoder's parameters join the flat
        arena (train_vpd_model.py:61-65). Draws its init from the global RNG now,
        like the reference does when it constructs FCNet."""
        if self._motion:
            return
        dsd = _init.decoder_state(self.emb_dim)
        old = self._params
        self._motion = True
        self._layout_net = _NativeNet(self.model_arch, self.emb_dim, self._in_channels, 128, 128,
                                      1, True)
        self._alloc_arenas(keep=old)
        self._net = None
        self._write_state(OrderedDict(('decoder.' + k, v) for k, v in dsd.items()), strict=False)

    def _native(self, height, width, batch):
        """The bound native net able to run `batch` frames of height x width."""
        need_grads = self._grads is not None
        n = self._net
        if (n is None or n.key[0] != height or n.key[1] != width or n.max_batch < batch
                or n.key[3] != self._motion or n.has_grads != need_grads):
            cap = batch if n is None or n.key[:2] != (height, width) else max(batch, n.max_batch)
            if n is not None:
                n.close()
            with torch.cuda.device(self._dev):
                n = _NativeNet(self.model_arch, self.emb_dim, self._in_channels, height, width,
                               cap, self._motion)
                nbytes = lib().call('vpd_net_workspace_bytes', n.handle)
                n.workspace = torch.empty(nbytes + 2048, device=self._dev, dtype=torch.uint8)
                lib().call('vpd_net_bind', n.handle, self._params, self._grads, self._buffers,
                           self._nbt, n.workspace, n.workspace.numel())
            n.has_grads = need_grads
            self._net = n
        return n

    # --------------------------------------------------------------- state dict
    def _views(self):
        """name -> (arena tensor view, layout, shape) for every reference tensor."""
        arenas = (self._params, self._buffers, self._nbt)
        out = OrderedDict()
        for name, arena, off, layout, shape in self._table:
            numel = int(np.prod(shape)) if shape else 1
            if layout == 1:
                cout, cin, kh, kw = shape
                v = arenas[arena][off:off + numel].view(kh * kw, cout, cin)
            elif layout == 2:
                v = arenas[arena][off:off + 7 * 64 * 64].view(7, 64, 8, 8)   # kh, co, kw, c
            else:
                v = arenas[arena][off:off + numel].view(shape)
            out[name] = (v, layout, shape)
        return out

    def _write_state(self, sd, strict=True):
        views = self._views()
        own = [k for k in views if not k.startswith('decoder.')]
        if strict:
            missing = [k for k in own if k not in sd]
            unexpected = [k for k in sd if k not in views or k.startswith('decoder.')]
            if missing or unexpected:
                raise RuntimeError('Error(s) in loading state_dict: missing {} unexpected {}'.format(
                    missing, unexpected))
        for k, t in sd.items():
            if k not in views:
                continue
            v, layout, shape = views[k]
            t = torch.as_tensor(t)
            if tuple(t.shape) != tuple(shape):
                raise RuntimeError('size mismatch for {}: {} vs {}'.format(
                    k, tuple(t.shape), tuple(shape)))
            t = t.to(self._dev)
            if layout == 1:
                cout, cin, kh, kw = shape
                v.copy_(t.float().permute(2, 3, 0, 1).reshape(kh * kw, cout, cin))
            elif layout == 2:
                v.zero_()
                v[:, :, :7, :shape[1]] = t.float().permute(2, 0, 3, 1)       # kh, co, kw, c
            else:
                v.copy_(t.to(v.dtype))
        if self._net is not None:
            lib().call('vpd_net_params_changed', self._net.handle)

    def _read_state(self, prefix_filter):
        out = OrderedDict()
        for k, (v, layout, shape) in self._views().items():
            if not prefix_filter(k):
                continue
            if layout == 1:
                cout, cin, kh, kw = shape
                t = v.view(kh, kw, cout, cin).permute(2, 3, 0, 1).contiguous()
            elif layout == 2:
                t = v[:, :, :7, :shape[1]].permute(1, 3, 0, 2).contiguous()  # co, c, kh, kw
            else:
                t = v.clone()
            out[k] = t
        return out

    def state_dict(self):
        """The reference's encoder state_dict (218 `resnet.*` entries for resnet34),
        OIHW fp32 conv weights, int64 num_batches_tracked; tensors on the device."""
        return self._read_state(lambda k: k.startswith('resnet.'))

    def load_state_dict(self, state_dict, strict=True):
        self._write_state(state_dict, strict=strict)

    def decoder_state_dict(self):
        """FCNet state ('layers.{0,2,5}.{weight,bias}'), train_vpd_model.py:110-112."""
        return OrderedDict((k[len('decoder.'):], v) for k, v in
                           self._read_state(lambda k: k.startswith('decoder.')).items())

    def load_decoder_state_dict(self, sd):
        self._attach_decoder()
        self._write_state(OrderedDict(('decoder.' + k, v) for k, v in sd.items()), strict=False)

    # ------------------------------------------------------------ nn.Module-ish
    def parameters(self):
        """One flat fp32 tensor holding every parameter (AdamW is element-wise, so
        any torch optimizer over it is valid); `.grad` is the gradient arena."""
        p = self._params
        p.grad = self._ensure_grads()
        return [p]

    def to(self, device):
        if torch.device('cuda' if str(device) == 'cuda' else device).type != 'cuda':
            raise VpdError('vpd_b200 models live on CUDA devices only')
        return self

    def train(self, mode=True):
        self.training = mode
        return self

    def eval(self):
        return self.train(False)

    # ---------------------------------------------------------------- forward
    def _check_channels(self, x):
        if self.use_flow:
            assert x.shape[1] == 5, 'Wrong number of channels for RGB + flow'
        else:
            assert x.shape[1] == 3, 'Wrong number of channels for RGB'

    def forward(self, x):
        """Encoder output [B, emb_dim] as a device tensor (models/rgb.py:68-70). In eval mode
        BatchNorm uses the running statistics; in train mode (the nn.Module default) the batch
        statistics, and the running buffers / counters are updated, like the reference module
        does - the whole batch in one launch group, because the statistics are per batch. There
        is no autograd graph either way: gradients come from ModelTrainer.epoch, which runs the
        native forward + backward."""
        x = x.to(self._dev, dtype=torch.float32).contiguous()
        self._check_channels(x)
        B, _, H, W = x.shape
        out = torch.empty((B, self.emb_dim), device=self._dev, dtype=torch.float32)
        with torch.cuda.device(self._dev):
            if self.training:
                net = self._native(H, W, B)
                lib().call('vpd_net_forward_train', net.handle, x, None, B, out,
                           stream_ptr(self._dev))
                return out
            for i in range(0, B, _MAX_CHUNK):
                n = min(_MAX_CHUNK, B - i)
                net = self._native(H, W, n)
                lib().call('vpd_net_forward', net.handle, x[i:i + n], None, n, out[i:i + n],
                           stream_ptr(self._dev))
        return out

    __call__ = forward

    def embed(self, x):
        """models/rgb.py:72-86: ndarray/Tensor [C,H,W] or [N,C,H,W] -> np.float32 [N, emb_dim]."""
        if not isinstance(x, torch.Tensor):
            # the reference's torch.Tensor(x): float32 values; as_tensor shares a float32
            # ndarray's memory instead of copying it first (10 MB at batch 32)
            x = torch.as_tensor(x, dtype=torch.float32)
        x = x.to(self._dev)
        if len(x.shape) == 3:
            x = x.unsqueeze(0)
        self._check_channels(x)
        self.eval()
        with torch.no_grad():
            return self(x).cpu().numpy()

    def embed_stem(self, stem, batch, height, width):
        """Embeddings for frames already in the network's input layout (written by
        vpd_assemble_stem): device tensor [batch, emb_dim]."""
        out = torch.empty((batch, self.emb_dim), device=self._dev, dtype=torch.float32)
        with torch.cuda.device(self._dev):
            net = self._native(height, width, batch)
            lib().call('vpd_net_forward', net.handle, None, stem, batch, out,
                       stream_ptr(self._dev))
        return out

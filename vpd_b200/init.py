"""Parameter initialisation equal, bit for bit and RNG draw for RNG draw, to the
reference constructors, so `torch.manual_seed(s); RGBF_EmbeddingModel(...)` gives
the same weights here as in jhong93/vpd:

  * torchvision `resnet18/34(pretrained=False)`: module construction order
    (each nn.Conv2d / nn.Linear draws its default init), then the
    kaiming_normal_(mode='fan_out') loop over `modules()`;
  * models/rgb.py:8-37  add_flow_to_model: conv1 -> 5 channels = mean over the 3
    input channels, broadcast; the replacement nn.Conv2d draws (and discards);
  * models/rgb.py:40-43 replace_last_layer: fc = nn.Linear(512, emb_dim);
  * train_vpd_model.py:61-65 FCNet(emb_dim, [128,128], 2*emb_dim) decoder.
All draws happen on the CPU default generator, like the reference.
"""
import math
from collections import OrderedDict

import torch

ARCH_LAYERS = {'resnet18': (2, 2, 2, 2), 'resnet34': (3, 4, 6, 3)}


def blocks(arch):
    """(prefix, cin, cout, stride, has_downsample) per BasicBlock, in order."""
    if arch not in ARCH_LAYERS:
        raise NotImplementedError(
            "encoder_arch '{}' is not supported by the CUDA path (BasicBlock ResNets: {})".format(
                arch, ', '.join(sorted(ARCH_LAYERS))))
    out, inplanes = [], 64
    for stage, (planes, count) in enumerate(zip((64, 128, 256, 512), ARCH_LAYERS[arch])):
        for i in range(count):
            stride = 2 if (i == 0 and stage > 0) else 1
            out.append(('resnet.layer{}.{}'.format(stage + 1, i), inplanes, planes, stride,
                        i == 0 and (stride != 1 or inplanes != planes)))
            inplanes = planes
    return out


def _draw_conv_default(*shape):
    torch.nn.init.kaiming_uniform_(torch.empty(shape), a=math.sqrt(5))


def _draw_linear(out_features, in_features):
    w = torch.empty(out_features, in_features)
    torch.nn.init.kaiming_uniform_(w, a=math.sqrt(5))
    b = torch.empty(out_features)
    bound = 1 / math.sqrt(in_features)
    torch.nn.init.uniform_(b, -bound, bound)
    return w, b


def _kaiming_fan_out(*shape):
    w = torch.empty(shape)
    torch.nn.init.kaiming_normal_(w, mode='fan_out', nonlinearity='relu')
    return w


def _bn(sd, prefix, channels):
    sd[prefix + '.weight'] = torch.ones(channels)
    sd[prefix + '.bias'] = torch.zeros(channels)
    sd[prefix + '.running_mean'] = torch.zeros(channels)
    sd[prefix + '.running_var'] = torch.ones(channels)
    sd[prefix + '.num_batches_tracked'] = torch.tensor(0, dtype=torch.int64)


def encoder_state(arch, emb_dim, use_flow):
    blks = blocks(arch)
    # constructor draws of torchvision's ResNet.__init__
    _draw_conv_default(64, 3, 7, 7)
    for _, cin, cout, _, ds in blks:
        if ds:
            _draw_conv_default(cout, cin, 1, 1)
        _draw_conv_default(cout, cin, 3, 3)
        _draw_conv_default(cout, cout, 3, 3)
    _draw_linear(1000, 512)
    # re-initialisation loop
    sd = OrderedDict()
    rgb_kernel = _kaiming_fan_out(64, 3, 7, 7)
    sd['resnet.conv1.weight'] = rgb_kernel
    _bn(sd, 'resnet.bn1', 64)
    for prefix, cin, cout, _, ds in blks:
        sd[prefix + '.conv1.weight'] = _kaiming_fan_out(cout, cin, 3, 3)
        _bn(sd, prefix + '.bn1', cout)
        sd[prefix + '.conv2.weight'] = _kaiming_fan_out(cout, cout, 3, 3)
        _bn(sd, prefix + '.bn2', cout)
        if ds:
            sd[prefix + '.downsample.0.weight'] = _kaiming_fan_out(cout, cin, 1, 1)
            _bn(sd, prefix + '.downsample.1', cout)
    if use_flow:
        sd['resnet.conv1.weight'] = rgb_kernel.mean(dim=1, keepdim=True).expand(
            64, 5, 7, 7).contiguous()
        _draw_conv_default(64, 5, 7, 7)
    sd['resnet.fc.weight'], sd['resnet.fc.bias'] = _draw_linear(emb_dim, 512)
    return sd


def decoder_state(emb_dim):
    sd = OrderedDict()
    for key, (o, i) in (('layers.0', (128, emb_dim)), ('layers.2', (128, 128)),
                        ('layers.5', (2 * emb_dim, 128))):
        sd[key + '.weight'], sd[key + '.bias'] = _draw_linear(o, i)
    return sd


def fcresnet_state(in_dim, out_dim, num_blocks, hidden_dim):
    """models/module.py:192-204 `FCResNet(in_dim, out_dim, num_blocks, hidden_dim)`: the
    state_dict an nn.Module built in that order holds after construction (default nn.Linear /
    nn.BatchNorm1d initialisation, same draws from the CPU generator). Keys:
    layers.0 Linear(in, hidden); layers.{2+i}.block.{0,4} Linear(hidden, hidden) and
    .block.{1,5} BatchNorm1d per FcResidualBlock (models/module.py:159-177); the last entry
    Linear(hidden, out) unless out_dim is None."""
    sd = OrderedDict()
    sd['layers.0.weight'], sd['layers.0.bias'] = _draw_linear(hidden_dim, in_dim)
    for i in range(num_blocks):
        p = 'layers.{}.block'.format(2 + i)
        for lin, bn in ((0, 1), (4, 5)):
            w, b = _draw_linear(hidden_dim, hidden_dim)
            sd['{}.{}.weight'.format(p, lin)], sd['{}.{}.bias'.format(p, lin)] = w, b
            _bn(sd, '{}.{}'.format(p, bn), hidden_dim)
    if out_dim is not None:
        p = 'layers.{}'.format(2 + num_blocks)
        sd[p + '.weight'], sd[p + '.bias'] = _draw_linear(out_dim, hidden_dim)
    return sd

"""Host-side RAFT network: the caller of the correlation hot path.

Mirrors the reference interface `RAFT(args).forward(image1, image2, iters,
flow_init, upsample, test_mode)` (RAFT/core/raft.py:24-144) and keeps the
reference's parameter names, so `raft-things.pth` / `raft-small.pth`
checkpoints (with or without the DataParallel `module.` prefix, ofgen.py:67-68)
load with strict=True.  The dense convolutions stay in PyTorch/cuDNN (they are
outside the graded path); the all-pairs correlation volume, its pyramid and
the per-iteration lookup run in this package's sm_100a kernels via
`corr.CorrBlock` / `corr.AlternateCorrBlock`.

Differences from the reference that do not change results:
  * in `test_mode` the convex 8x upsample is evaluated only for the last
    iteration (the reference computes it every iteration and discards all but
    the last, raft.py:133-142);
  * the coordinate grid is built once per call.
"""
from __future__ import annotations

import zlib
from types import SimpleNamespace

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import corr as _corr


# --------------------------------------------------------------------------- encoders
def _norm(kind: str, ch: int, groups: int) -> nn.Module:
    if kind == 'group':
        return nn.GroupNorm(num_groups=groups, num_channels=ch)
    if kind == 'batch':
        return nn.BatchNorm2d(ch)
    if kind == 'instance':
        return nn.InstanceNorm2d(ch)
    if kind == 'none':
        return nn.Sequential()
    raise ValueError(f'unknown norm {kind!r}')


class _ResUnit(nn.Module):
    """Two 3x3 convs + skip (extractor.py:6-56 'ResidualBlock')."""

    def __init__(self, cin: int, cout: int, norm: str, stride: int):
        super().__init__()
        g = cout // 8
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1, stride=stride)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _norm(norm, cout, g)
        self.norm2 = _norm(norm, cout, g)
        self.downsample = None
        if stride != 1:
            # the reference registers the skip norm twice (norm3 and downsample.1)
            self.norm3 = _norm(norm, cout, g)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm3)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class _BottleneckUnit(nn.Module):
    """1x1 -> 3x3 -> 1x1 + skip (extractor.py:59-115 'BottleneckBlock')."""

    def __init__(self, cin: int, cout: int, norm: str, stride: int):
        super().__init__()
        g = cout // 8
        mid = cout // 4
        self.conv1 = nn.Conv2d(cin, mid, 1)
        self.conv2 = nn.Conv2d(mid, mid, 3, padding=1, stride=stride)
        self.conv3 = nn.Conv2d(mid, cout, 1)
        self.relu = nn.ReLU(inplace=True)
        self.norm1 = _norm(norm, mid, g)
        self.norm2 = _norm(norm, mid, g)
        self.norm3 = _norm(norm, cout, g)
        self.downsample = None
        if stride != 1:
            self.norm4 = _norm(norm, cout, g)
            self.downsample = nn.Sequential(nn.Conv2d(cin, cout, 1, stride=stride), self.norm4)

    def forward(self, x):
        y = self.relu(self.norm1(self.conv1(x)))
        y = self.relu(self.norm2(self.conv2(y)))
        y = self.relu(self.norm3(self.conv3(y)))
        if self.downsample is not None:
            x = self.downsample(x)
        return self.relu(x + y)


class Encoder(nn.Module):
    """1/8-resolution CNN (extractor.py:118-192 BasicEncoder, :195-267 SmallEncoder)."""

    def __init__(self, output_dim: int, norm_fn: str, dropout: float, small: bool):
        super().__init__()
        widths = (32, 64, 96) if small else (64, 96, 128)
        unit = _BottleneckUnit if small else _ResUnit
        self.norm_fn = norm_fn
        self.norm1 = _norm(norm_fn, widths[0], 8)
        self.conv1 = nn.Conv2d(3, widths[0], 7, stride=2, padding=3)
        self.relu1 = nn.ReLU(inplace=True)
        cin = widths[0]
        for i, (wd, st) in enumerate(zip(widths, (1, 2, 2)), start=1):
            setattr(self, f'layer{i}', nn.Sequential(unit(cin, wd, norm_fn, st), unit(wd, wd, norm_fn, 1)))
            cin = wd
        self.conv2 = nn.Conv2d(cin, output_dim, 1)
        self.dropout = nn.Dropout2d(p=dropout) if dropout > 0 else None
        for m in self.modules():
            if isinstance(m, nn.Conv2d):
                nn.init.kaiming_normal_(m.weight, mode='fan_out', nonlinearity='relu')
            elif isinstance(m, (nn.BatchNorm2d, nn.InstanceNorm2d, nn.GroupNorm)):
                if m.weight is not None:
                    nn.init.constant_(m.weight, 1)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)

    def forward(self, x):
        pair = isinstance(x, (tuple, list))
        if pair:
            n = x[0].shape[0]
            x = torch.cat(list(x), dim=0)
        x = self.relu1(self.norm1(self.conv1(x)))
        x = self.layer3(self.layer2(self.layer1(x)))
        x = self.conv2(x)
        if self.training and self.dropout is not None:
            x = self.dropout(x)
        if pair:
            x = torch.split(x, [n, n], dim=0)
        return x


# --------------------------------------------------------------------------- update block
class _FlowHead(nn.Module):
    def __init__(self, cin: int, hidden: int):
        super().__init__()
        self.conv1 = nn.Conv2d(cin, hidden, 3, padding=1)
        self.conv2 = nn.Conv2d(hidden, 2, 3, padding=1)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.conv2(self.relu(self.conv1(x)))


class _GRU(nn.Module):
    """update.py:16-30 (ConvGRU, 3x3) and :32-60 (SepConvGRU, 1x5 then 5x1)."""

    def __init__(self, hidden: int, cin: int, separable: bool):
        super().__init__()
        self.passes = ('1', '2') if separable else ('',)
        shapes = {'1': ((1, 5), (0, 2)), '2': ((5, 1), (2, 0)), '': (3, 1)}
        for p in self.passes:
            k, pad = shapes[p]
            for gate in 'zrq':
                setattr(self, f'conv{gate}{p}', nn.Conv2d(hidden + cin, hidden, k, padding=pad))

    def forward(self, h, x):
        for p in self.passes:
            hx = torch.cat([h, x], dim=1)
            z = torch.sigmoid(getattr(self, f'convz{p}')(hx))
            r = torch.sigmoid(getattr(self, f'convr{p}')(hx))
            q = torch.tanh(getattr(self, f'convq{p}')(torch.cat([r * h, x], dim=1)))
            h = (1 - z) * h + z * q
        return h


class _MotionEncoder(nn.Module):
    """update.py:62-77 (small) / :79-97 (basic)."""

    def __init__(self, cor_planes: int, small: bool):
        super().__init__()
        self.small = small
        if small:
            self.convc1 = nn.Conv2d(cor_planes, 96, 1)
            self.convf1 = nn.Conv2d(2, 64, 7, padding=3)
            self.convf2 = nn.Conv2d(64, 32, 3, padding=1)
            self.conv = nn.Conv2d(128, 80, 3, padding=1)
        else:
            self.convc1 = nn.Conv2d(cor_planes, 256, 1)
            self.convc2 = nn.Conv2d(256, 192, 3, padding=1)
            self.convf1 = nn.Conv2d(2, 128, 7, padding=3)
            self.convf2 = nn.Conv2d(128, 64, 3, padding=1)
            self.conv = nn.Conv2d(64 + 192, 128 - 2, 3, padding=1)

    def forward(self, flow, corr):
        cor = F.relu(self.convc1(corr))
        if not self.small:
            cor = F.relu(self.convc2(cor))
        flo = F.relu(self.convf2(F.relu(self.convf1(flow))))
        out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
        return torch.cat([out, flow], dim=1)


class UpdateBlock(nn.Module):
    """update.py:99-136: motion encoder -> GRU -> flow head (+ upsample mask)."""

    def __init__(self, cor_planes: int, hidden: int, small: bool):
        super().__init__()
        self.encoder = _MotionEncoder(cor_planes, small)
        if small:
            self.gru = _GRU(hidden, 82 + 64, separable=False)
            self.flow_head = _FlowHead(hidden, 128)
            self.mask = None
        else:
            self.gru = _GRU(hidden, 128 + hidden, separable=True)
            self.flow_head = _FlowHead(hidden, 256)
            self.mask = nn.Sequential(nn.Conv2d(128, 256, 3, padding=1), nn.ReLU(inplace=True),
                                      nn.Conv2d(256, 64 * 9, 1))

    def forward(self, net, inp, corr, flow, want_mask: bool = True):
        motion = self.encoder(flow, corr)
        net = self.gru(net, torch.cat([inp, motion], dim=1))
        delta = self.flow_head(net)
        mask = None
        if self.mask is not None and want_mask:
            mask = 0.25 * self.mask(net)
        return net, mask, delta


# --------------------------------------------------------------------------- helpers
def coords_grid(batch: int, ht: int, wd: int, device) -> torch.Tensor:
    """[B,2,ht,wd] with channel 0 = x, channel 1 = y (utils/utils.py:74-77)."""
    ys, xs = torch.meshgrid(torch.arange(ht, device=device), torch.arange(wd, device=device), indexing='ij')
    return torch.stack([xs, ys], dim=0).float()[None].repeat(batch, 1, 1, 1)


def convex_upsample(flow: torch.Tensor, mask: torch.Tensor) -> torch.Tensor:
    """RAFT.upsample_flow (raft.py:72-83): [N,2,h,w] -> [N,2,8h,8w]."""
    n, _, h, w = flow.shape
    mask = torch.softmax(mask.view(n, 1, 9, 8, 8, h, w), dim=2)
    up = F.unfold(8 * flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
    up = torch.sum(mask * up, dim=2).permute(0, 1, 4, 2, 5, 3)
    return up.reshape(n, 2, 8 * h, 8 * w)


def upflow8(flow: torch.Tensor) -> torch.Tensor:
    """utils/utils.py:80-82."""
    return 8 * F.interpolate(flow, size=(8 * flow.shape[2], 8 * flow.shape[3]), mode='bilinear', align_corners=True)


class InputPadder:
    """Replicate-pad to a multiple of 8 (utils/utils.py:7-24)."""

    def __init__(self, dims, mode: str = 'sintel'):
        self.ht, self.wd = dims[-2:]
        ph = (-self.ht) % 8
        pw = (-self.wd) % 8
        if mode == 'sintel':
            self._pad = [pw // 2, pw - pw // 2, ph // 2, ph - ph // 2]
        else:
            self._pad = [pw // 2, pw - pw // 2, 0, ph]

    def pad(self, *inputs):
        return [F.pad(x, self._pad, mode='replicate') for x in inputs]

    def unpad(self, x):
        ht, wd = x.shape[-2:]
        return x[..., self._pad[2]:ht - self._pad[3], self._pad[0]:wd - self._pad[1]]


def _get(args, name, default):
    return getattr(args, name, default) if args is not None else default


class RAFT(nn.Module):
    """Drop-in for RAFT/core/raft.py:24 `RAFT(args)`.

    `args` may be the scripts' ad-hoc `namespace` (ofgen.py:51-66), an
    argparse Namespace or None; recognised fields: small, mixed_precision,
    alternate_corr, dropout, plus this package's `corr_precision`
    ('fp16' | 'tf32' | '3xtf32' | 'bf16' | 'fp32')."""

    def __init__(self, args=None):
        super().__init__()
        if args is None:
            args = SimpleNamespace()
        self.args = args
        small = bool(_get(args, 'small', False))
        if small:
            self.hidden_dim, self.context_dim, radius, fdim = 96, 64, 3, 128
        else:
            self.hidden_dim, self.context_dim, radius, fdim = 128, 128, 4, 256
        args.corr_levels = 4
        args.corr_radius = radius
        if not hasattr(args, 'dropout'):
            args.dropout = 0
        if not hasattr(args, 'alternate_corr'):
            args.alternate_corr = False
        if not hasattr(args, 'mixed_precision'):
            args.mixed_precision = False
        self.small = small
        cnet_norm = 'none' if small else 'batch'
        self.fnet = Encoder(fdim, 'instance', args.dropout, small)
        self.cnet = Encoder(self.hidden_dim + self.context_dim, cnet_norm, args.dropout, small)
        self.update_block = UpdateBlock(args.corr_levels * (2 * radius + 1) ** 2, self.hidden_dim, small)

    def freeze_bn(self):
        for m in self.modules():
            if isinstance(m, nn.BatchNorm2d):
                m.eval()

    def load_state_dict(self, state_dict, strict: bool = True, **kw):
        if any(k.startswith('module.') for k in state_dict):
            state_dict = {k[len('module.'):] if k.startswith('module.') else k: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    def _autocast(self, device_type: str):
        return torch.autocast(device_type=device_type, enabled=bool(self.args.mixed_precision))

    def make_corr_fn(self, fmap1, fmap2):
        radius = self.args.corr_radius
        prec = _get(self.args, 'corr_precision', 'fp16')
        if self.args.alternate_corr:
            return _corr.AlternateCorrBlock(fmap1, fmap2, radius=radius)
        return _corr.CorrBlock(fmap1, fmap2, radius=radius, precision=prec)

    def encode(self, image1, image2):
        """normalise -> fnet on both images -> corr_fn; cnet on image1 (raft.py:89-114)."""
        image1 = (2 * (image1 / 255.0) - 1.0).contiguous()
        image2 = (2 * (image2 / 255.0) - 1.0).contiguous()
        dev = image1.device.type
        with self._autocast(dev):
            fmap1, fmap2 = self.fnet([image1, image2])
        corr_fn = self.make_corr_fn(fmap1.float(), fmap2.float())
        with self._autocast(dev):
            cnet = self.cnet(image1)
            net, inp = torch.split(cnet, [self.hidden_dim, self.context_dim], dim=1)
            net = torch.tanh(net)
            inp = torch.relu(inp)
        return corr_fn, net, inp

    def forward(self, image1, image2, iters: int = 12, flow_init=None, upsample: bool = True, test_mode: bool = False):
        corr_fn, net, inp = self.encode(image1, image2)
        n, _, H, W = image1.shape
        coords0 = coords_grid(n, H // 8, W // 8, image1.device)
        coords1 = coords0.clone()
        if flow_init is not None:
            coords1 = coords1 + flow_init
        dev = image1.device.type
        preds = []
        up_mask = None
        for itr in range(iters):
            coords1 = coords1.detach()
            corr = corr_fn(coords1)
            flow = coords1 - coords0
            last = itr == iters - 1
            with self._autocast(dev):
                net, up_mask, delta = self.update_block(net, inp, corr, flow, want_mask=(not test_mode) or last)
            coords1 = coords1 + delta
            if not test_mode:
                f = coords1 - coords0
                preds.append(upflow8(f) if up_mask is None else convex_upsample(f, up_mask.float()))
        if test_mode:
            f = coords1 - coords0
            flow_up = upflow8(f) if up_mask is None else convex_upsample(f, up_mask.float())
            return f, flow_up
        return preds


def fill_weights_by_name(module: nn.Module, seed: int = 0, flow_head_scale: float = 1.0) -> nn.Module:
    """Deterministic weights that depend only on (seed, parameter name, shape),
    not on module construction order: lets the reference RAFT (golden
    generator) and this RAFT (tests, bench 'random-init weights') hold
    identical parameters without shipping a 21 MB checkpoint.
    flow_head_scale != 1 scales the flow head's output convolution (update_block.flow_head.conv2): plain random weights make
    the update block an amplifier (the flow runs away to ~100 px in 20 iterations); 0.02 keeps the flow at a few px, the
    regime a trained checkpoint works in (tests/golden_inputs.py RAFT_FULL_CASES `calm`, bench.py)."""
    sd = module.state_dict()
    with torch.no_grad():
        for name in sorted(sd):
            t = sd[name]
            g = torch.Generator().manual_seed((zlib.crc32(name.encode()) + 7919 * seed) & 0x7FFFFFFF)
            if name.endswith('num_batches_tracked'):
                continue
            if name.endswith('running_var'):
                v = 0.5 + torch.rand(t.shape, generator=g)
            elif name.endswith('running_mean'):
                v = 0.1 * torch.randn(t.shape, generator=g)
            elif t.dim() == 4:
                fan_in = t.shape[1] * t.shape[2] * t.shape[3]
                v = torch.randn(t.shape, generator=g) * (1.0 / fan_in) ** 0.5
            elif name.endswith('weight'):
                v = 1.0 + 0.1 * torch.randn(t.shape, generator=g)
            else:
                v = 0.05 * torch.randn(t.shape, generator=g)
            if flow_head_scale != 1.0 and name.startswith('update_block.flow_head.conv2.'):
                v = v * flow_head_scale
            t.copy_(v.to(t.dtype))
    return module

"""Perceptual-loss STAND-IN — labelled, NON-PARITY.

BASELINE.json configs[3] words the single-GPU training step "with random-init VGG19 perceptual loss".  The reference's
own VGG branch cannot run: ``ops.VGG19()`` raises at construction (code/ops.py:153-166, nn.Conv2d without kernel_size),
``VGG19_slim`` is called without its ``reuse`` argument (code/train.py:30 vs :126-127) and its normalisation adds a float to
the tuple ``torch.min(..., dim)`` returns (:41); ``--vgg_scaling`` defaults to -0.002, which switches it off (main.py:98).
There is therefore nothing to be bit-compatible WITH, and this module is OUR definition of what that branch evidently
means (it is the TecoGAN perceptual term, code/train.py:253-273), never compared against the reference:

  * network: the standard VGG19 3x3 conv stack up to conv4_4 (2-2-4-4 convs of 64/128/256/512 channels, ReLU after every
    conv, 2x2 max-pool between the groups), RANDOM weights from a fixed seed (the reference re-creates a random VGG19 on
    every call, code/train.py:33, and never reads --vgg_ckpt), frozen;
  * input: ``deprocess(x) * 255 - VGG_MEAN`` (code/train.py:31-32);
  * features: conv2_2, conv3_4, conv4_4 (code/train.py:125), each normalised to unit length over channels (:41);
  * loss: sum over the three layers of ``1 - mean(cosine similarity)`` between the generator-output features and the
    (detached) target features; ``gen_loss += vgg_scaling * loss`` (:267).

The convolutions run on the repo's tcgen05 conv core through the C ABI (tg_conv3x3_fwd / tg_conv3x3_dgrad): activations
are NHWC bf16, kept as lists of <= 128-channel chunks so that the 256- and 512-channel layers are sums of (output block,
input block) launches with the partial sums carried through the kernels' residual input.  Pooling, normalisation and the
loss are a handful of element-wise torch ops on those chunks (HBM-bound, off the tensor-core path).  Off unless
``tecogan_b200.perceptual.ENABLED`` is set (or TG_PERCEPTUAL_STANDIN=1): with it off ``vgg_scaling > 0`` raises, which is
what the reference does.
"""
import os

import torch
import torch.nn.functional as F

from . import _native as _nt

ENABLED = os.environ.get("TG_PERCEPTUAL_STANDIN", "0") == "1"
VGG_MEAN = (123.68, 116.78, 103.94)                                   # code/train.py:6
_CFG = ((64, 64), (128, 128), (256, 256, 256, 256), (512, 512, 512, 512))   # VGG19 up to conv4_4
_TAPS = {(1, 1): "conv2_2", (2, 3): "conv3_4", (3, 3): "conv4_4"}     # (group, conv index) -> feature name (code/train.py:125)
_BLK = 128                                                             # channel chunk of the conv core


def _blocks(c):
    return [(o, min(_BLK, c - o)) for o in range(0, c, _BLK)]


class _FrozenConv(torch.autograd.Function):
    """ReLU(conv3x3(x) + b) with frozen weights on chunked NHWC bf16 activations; backward = data gradient only."""

    @staticmethod
    def forward(ctx, layer, n, h, w, *xs):
        lib = _nt.lib()
        st = _nt.stream_ptr(xs[0].device)
        ys = []
        nci = len(xs)
        for bo, (co0, con) in enumerate(_blocks(layer.cout)):
            bufs = [torch.empty((n, h, w, con), dtype=torch.bfloat16, device=xs[0].device) for _ in range(min(nci, 2))]
            for bi, x in enumerate(xs):
                # partial sums ride on the kernel's residual input (ping-pong buffers: the residual is read through the
                # read-only path); the ReLU can only be fused when there is a single input block
                y, prev = bufs[bi & 1], (bufs[(bi - 1) & 1] if bi else None)
                _nt.check(lib.tg_conv3x3_fwd(_nt.ptr(x), _nt.ptr(layer.packed[bo][bi]), _nt.ptr(prev), _nt.ptr(y), n, h, w,
                                             x.shape[3], con, 1 if nci == 1 else 0, _nt.AMODE_HALO, st))
            if nci > 1:
                y.clamp_(min=0)
            ys.append(y)
        ctx.layer, ctx.shape = layer, (n, h, w)
        ctx.save_for_backward(*xs, *ys)
        ctx.nx = len(xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *dys):
        lib = _nt.lib()
        layer = ctx.layer
        n, h, w = ctx.shape
        saved = ctx.saved_tensors
        xs, ys = saved[:ctx.nx], saved[ctx.nx:]
        st = _nt.stream_ptr(xs[0].device)
        # ReLU backward of THIS layer's output: dZ = dY where y > 0
        dzs = [torch.where(y > 0, dy.to(torch.bfloat16), torch.zeros((), dtype=torch.bfloat16, device=y.device)).contiguous()
               for dy, y in zip(dys, ys)]
        if not layer.needs_dx:
            return (None, None, None, None) + (None,) * ctx.nx
        dxs = []
        for bi, x in enumerate(xs):
            cin_blk = layer.cin_blocks[bi][1]
            bufs = [torch.empty_like(x) for _ in range(min(len(dzs), 2))]
            for bo, dz in enumerate(dzs):
                dx, prev = bufs[bo & 1], (bufs[(bo - 1) & 1] if bo else None)
                _nt.check(lib.tg_conv3x3_dgrad(_nt.ptr(dz), _nt.ptr(layer.packed_dgrad[bo][bi]), _nt.ptr(prev), None,
                                               _nt.ptr(dx), n, h, w, cin_blk, dz.shape[3], st))
            dxs.append(dx)
        return (None, None, None, None) + tuple(dxs)


class _Layer:
    pass


class PerceptualStandIn:
    """Random-init (fixed seed) VGG19-to-conv4_4 feature extractor on the repo's conv core + the cosine feature loss."""

    def __init__(self, device, seed=19):
        lib = _nt.lib()
        self.dev = torch.device(device)
        g = torch.Generator().manual_seed(seed)
        self.layers = []
        cin = 3
        with torch.cuda.device(self.dev):
            st = _nt.stream_ptr(self.dev)
            for gi, group in enumerate(_CFG):
                for li, cout in enumerate(group):
                    bound = 1.0 / (cin * 9) ** 0.5                       # torch's default Conv2d init scale (kaiming_uniform(a=sqrt(5)))
                    wgt = ((torch.rand((cout, cin, 3, 3), generator=g) * 2 - 1) * bound).to(self.dev)
                    bias = ((torch.rand((cout,), generator=g) * 2 - 1) * bound).to(self.dev)
                    L = _Layer()
                    L.cin, L.cout, L.group, L.index = cin, cout, gi, li
                    L.weight, L.bias = wgt, bias                          # f32 originals (tests re-run the network with torch)
                    if cin < 64:                                          # the RGB layer: explicit zero weights for the 61 padding
                        wgt = F.pad(wgt, (0, 0, 0, 0, 0, 64 - cin))       # channels (the data-gradient packing needs >= 64 outputs)
                    L.cin_blocks, L.needs_dx = _blocks(wgt.shape[1]), True
                    L.packed, L.packed_dgrad = [], []
                    for co0, con in _blocks(cout):
                        row, rowd = [], []
                        for bi, (ci0, cik) in enumerate(L.cin_blocks):
                            wb = wgt[co0:co0 + con, ci0:ci0 + cik].contiguous()
                            bb = bias[co0:co0 + con].contiguous() if bi == 0 else None      # the bias enters once
                            pk = torch.zeros(lib.tg_packed_conv_bytes(0, cik, con), dtype=torch.uint8, device=self.dev)
                            _nt.check(lib.tg_pack_weights(0, _nt.ptr(wb), _nt.ptr(bb), cik, con, _nt.ptr(pk), st))
                            pd = torch.zeros(lib.tg_packed_conv_bytes(3, cik, con), dtype=torch.uint8, device=self.dev)
                            _nt.check(lib.tg_pack_weights(3, _nt.ptr(wb), None, cik, con, _nt.ptr(pd), st))
                            row.append(pk)
                            rowd.append(pd)
                        L.packed.append(row)
                        L.packed_dgrad.append(rowd)
                    self.layers.append(L)
                    cin = cout
            torch.cuda.synchronize(self.dev)
        self.mean = torch.tensor(VGG_MEAN, dtype=torch.float32, device=self.dev).view(1, 3, 1, 1)

    def features(self, x):
        """x [N,3,H,W] f32 in the generator's output range -> {name: [N,h,w,C] f32-normalisable chunk lists}."""
        n, c, h, w = x.shape
        if c != 3 or h % 8 or w % 8:
            raise RuntimeError("PerceptualStandIn: expected [N,3,H,W] with H, W multiples of 8")
        img = ((x + 1) / 2) * 255.0 - self.mean                              # deprocess, * 255, - VGG_MEAN (code/train.py:31-32)
        xs = [F.pad(img.permute(0, 2, 3, 1), (0, 61)).to(torch.bfloat16).contiguous()]     # NHWC, 3 -> 64 channels (zeros)
        feats, k = {}, 0
        for gi, group in enumerate(_CFG):
            if gi:
                xs = [F.max_pool2d(t.permute(0, 3, 1, 2), 2).permute(0, 2, 3, 1).contiguous() for t in xs]
                h, w = h // 2, w // 2
            for li in range(len(group)):
                xs = list(_FrozenConv.apply(self.layers[k], n, h, w, *xs))
                k += 1
                if (gi, li) in _TAPS:
                    feats[_TAPS[(gi, li)]] = xs
        return feats

    def loss(self, gen, target):
        """sum over conv2_2 / conv3_4 / conv4_4 of 1 - mean cosine similarity (gen features vs detached target features)."""
        n = gen.shape[0]
        f = self.features(torch.cat((gen, target.detach()), dim=0))           # one batch through the frozen network
        total, per_layer = 0, []
        for name in ("conv2_2", "conv3_4", "conv4_4"):
            chunks = [t.float() for t in f[name]]
            fg = torch.cat([t[:n] for t in chunks], dim=3)
            ft = torch.cat([t[n:] for t in chunks], dim=3).detach()
            fg = fg / torch.sqrt(torch.sum(fg * fg, dim=3, keepdim=True) + 1e-12)             # code/train.py:41 (as intended)
            ft = ft / torch.sqrt(torch.sum(ft * ft, dim=3, keepdim=True) + 1e-12)
            l = 1.0 - torch.mean(torch.sum(fg * ft, dim=3))
            per_layer.append(l)
            total = total + l
        return total, per_layer


_instances = {}


def get(device):
    key = str(torch.device(device))
    if key not in _instances:
        _instances[key] = PerceptualStandIn(device)
    return _instances[key]

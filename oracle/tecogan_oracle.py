"""torch-CPU fp32 restatement of the reference hot path (TEST INFRASTRUCTURE, see
oracle/__init__.py).  kind = "port": the reference is pure Python over PyTorch, it cannot
travel to the GPU box, so its few lines of model/loop structure are restated here over the
same third-party library (torch.nn.functional on CPU).  Pinned to the real reference by
tests/golden/*.npz (oracle/make_golden.py).

Every function cites the /root/reference file:line it follows.
"""
import types

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# ----------------------------------------------------------------------------- ops.py
def preprocess(image):            # code/ops.py:24-26
    return image * 2 - 1


def deprocess(image):             # code/ops.py:29-31
    return (image + 1) / 2


def upscale_four(x):              # code/ops.py:98-100 (align_corners=False default)
    return F.interpolate(x, scale_factor=4, mode="bilinear", align_corners=False)


def warp(img, grid):
    """main.py:203 / code/train.py:98: F.grid_sample(img, grid.half()) with defaults
    (bilinear, zeros, align_corners=False).  CPU grid_sample rejects mixed dtypes, so the
    fp16 rounding is written .half().float()."""
    return F.grid_sample(img.float(), grid.half().float(), mode="bilinear",
                         padding_mode="zeros", align_corners=False)


def space_to_depth(x, r=4):       # main.py:207-212
    n, c, hh, ww = x.shape
    h, w = hh // r, ww // r
    return x.view(n, c, h, r, w, r).permute(0, 1, 3, 5, 2, 4).reshape(n, c * r * r, h, w)


def depth_to_space(x, r=4):       # inverse of the above (north_star; == F.pixel_shuffle)
    return F.pixel_shuffle(x, r)


def _conv2(cin, k, cout, stride=1, use_bias=True):           # code/ops.py:57-63
    return nn.Conv2d(cin, cout, k, stride, padding=int((k - 1) / 2), bias=use_bias)


def _conv2_tran(cin, k, cout, stride, output_padding):       # code/ops.py:45-54
    return nn.ConvTranspose2d(cin, cout, k, stride, padding=int((k - 1) / 2), bias=True,
                              output_padding=output_padding)


# --------------------------------------------------------------------------- models.py
def _residual_block(cin, cout):                              # code/models.py:54-58
    return nn.Sequential(_conv2(cin, 3, cout, 1, True), nn.ReLU(), _conv2(cout, 3, cout, 1, False))


class OracleGenerator(nn.Module):
    """code/models.py:61-86.  Same attribute names => same state_dict keys."""

    def __init__(self, gen_output_channels=3, num_resblock=16):
        super().__init__()
        self.conv = nn.Sequential(_conv2(51, 3, 64, 1), nn.ReLU())
        self.resids = nn.ModuleList([_residual_block(64, 64) for _ in range(int(num_resblock))])
        self.conv_trans = nn.Sequential(
            _conv2_tran(64, 3, 64, 2, 1), nn.ReLU(),
            _residual_block(64, 64), _residual_block(64, 128),       # NO skip here (:73)
            _conv2_tran(128, 3, 128, 2, 1), nn.ReLU(),
            _conv2(128, 3, 64, 1), nn.ReLU())
        self.output = _conv2(64, 3, gen_output_channels, 1)

    def features(self, x):
        """pre-sigmoid logits (extra probe, not in the reference)."""
        net = self.conv(x)
        for block in self.resids:                            # :81-82
            net = block(net) + net
        net = self.conv_trans(net)
        return self.output(net)

    def forward(self, x):                                    # :78-86
        return torch.sigmoid(self.features(x))


def _discriminator_block(cin, cout, k, stride):              # code/models.py:90-94
    return nn.Sequential(_conv2(cin, k, cout, stride, False),
                         nn.BatchNorm2d(cout, eps=0.001),    # code/ops.py:75-77
                         nn.LeakyReLU(0.2))


class OracleDiscriminator(nn.Module):
    """code/models.py:97-146.  fc in-features 48 hard-codes 32x32 LR crops (:123);
    ``fc_in`` generalises it the way colab/README.md:15-22 tells users to."""

    def __init__(self, discrim_resblocks=4, discrim_channels=128, fc_in=48):
        super().__init__()
        ch = discrim_channels
        self.conv = nn.Sequential(_conv2(27, 3, 64, 1), nn.LeakyReLU(0.2))
        self.block1 = _discriminator_block(64, 64, 4, 2)
        self.resids1 = nn.ModuleList([nn.Sequential(_residual_block(64, 64), nn.BatchNorm2d(64, eps=0.001))
                                      for _ in range(int(discrim_resblocks))])
        self.block2 = _discriminator_block(64, ch, 4, 2)
        self.resids2 = nn.ModuleList([nn.Sequential(_residual_block(ch, ch), nn.BatchNorm2d(ch, eps=0.001))
                                      for _ in range(int(discrim_resblocks))])
        self.block3 = _discriminator_block(ch, ch, 4, 2)
        self.resids3 = nn.ModuleList([nn.Sequential(_residual_block(ch, ch), nn.BatchNorm2d(ch, eps=0.001))
                                      for _ in range(int(discrim_resblocks))])
        self.block4 = _discriminator_block(ch, 64, 4, 2)
        self.block5 = _discriminator_block(64, 3, 4, 2)
        self.fc = nn.Linear(fc_in, 1)                        # code/ops.py:85-88

    def forward(self, x):                                    # :125-146
        layer_list = []
        net = self.conv(x)
        net = self.block1(net)
        for block in self.resids1:
            net = block(net) + net
        layer_list.append(net)
        net = self.block2(net)
        for block in self.resids2:
            net = block(net) + net
        layer_list.append(net)
        net = self.block3(net)
        for block in self.resids3:
            net = block(net) + net
        layer_list.append(net)
        net = self.block4(net)
        layer_list.append(net)
        net = self.block5(net)
        net = net.view(net.shape[0], -1)
        net = self.fc(net)
        return torch.sigmoid(net), layer_list


class _RoundBf16(torch.autograd.Function):
    """x -> bf16 -> f32 in the forward direction; the gradient is rounded the same way (or passed through)."""

    @staticmethod
    def forward(ctx, x, round_grad):
        ctx.round_grad = round_grad
        return x.to(torch.bfloat16).float()

    @staticmethod
    def backward(ctx, g):
        return (g.to(torch.bfloat16).float() if ctx.round_grad else g), None


def emulate_bf16_operands(module, round_grads=True):
    """Turn an oracle network into the fp32 model of what the bf16 tensor-core path computes: every convolution sees
    bf16-rounded weights and bf16-rounded input activations (and, in the backward direction, bf16-rounded gradient
    tensors), everything else — accumulation, bias, BatchNorm, activations, residual adds — stays fp32.  Used by the
    gradient parity tests to separate kernel errors from the rounding noise that any bf16 implementation of the same
    network has (the reference itself trains under fp16 autocast, code/train.py:68,335-342).  In place; returns module."""
    for m in module.modules():
        if isinstance(m, (nn.Conv2d, nn.ConvTranspose2d)):
            with torch.no_grad():
                m.weight.copy_(m.weight.to(torch.bfloat16).float())
            m.register_forward_pre_hook(lambda mod, inp: (_RoundBf16.apply(inp[0], round_grads),))
    return module


def load_numpy_state(module, named):
    """copy {name: np.ndarray} into module (strict)."""
    sd = {k: torch.from_numpy(np.array(v, copy=True)) for k, v in named.items()}
    module.load_state_dict(sd, strict=True)
    return module


# ------------------------------------------------------------------ main.py frame loop
@torch.no_grad()
def frame_input(lr_t, lr_prev, prev_hr):
    """main.py:199-213 for one step: cat(LR_t, s2d(deprocess(warp(HR_{t-1}, flow)))).
    flow = upscale_four(LR_{t-1}*4)[:,0:2] .view-ed as [N,Ho,Wo,2] (main.py:186-189,200-201)."""
    n, _, h, w = lr_t.shape
    flow = upscale_four(lr_prev * 4.0)[:, 0:2].contiguous().view(n, 4 * h, 4 * w, 2)
    wp = warp(prev_hr, flow)
    return torch.cat((lr_t, space_to_depth(deprocess(wp), 4)), dim=1)


@torch.no_grad()
def infer_clip(G, r_inputs, return_inputs=False):
    """Restatement of /root/reference/main.py:173-219 with .cuda()/.cpu() removed, (H,W)
    generalised from the square crop_size, batch generalised from 1, fp32 throughout.
    r_inputs [B,T,3,H,W] -> [B,T,3,4H,4W]."""
    b, t, c, h, w = r_inputs.shape
    frame_t_pre = r_inputs[:, 0:-1]                                         # :181
    fnet_input = torch.reshape(frame_t_pre, (b * (t - 1), c, h, w))         # :183-184
    gen_flow = upscale_four(fnet_input * 4.0)                               # :186
    gen_flow = torch.reshape(gen_flow[:, 0:2], (b, t - 1, 2, h * 4, w * 4))  # :188-189
    input0 = torch.cat((r_inputs[:, 0], torch.zeros(b, 48, h, w)), dim=1)   # :191-193
    outs, ins = [], [input0]
    prev = G(input0).view(b, 3, h * 4, w * 4)                               # :195-196
    outs.append(prev)
    for i in range(t - 1):                                                  # :199
        cur_flow = gen_flow[:, i].contiguous().view(b, h * 4, w * 4, 2)     # :200-201
        wp = warp(prev, cur_flow)                                           # :203
        wp = deprocess(wp)                                                  # :206
        s2d = space_to_depth(wp, 4)                                         # :207-212
        inputs = torch.cat((r_inputs[:, i + 1], s2d), dim=1)                # :213
        prev = G(inputs)                                                    # :214
        ins.append(inputs)
        outs.append(prev)
    out = torch.stack(outs, dim=1)                                          # :218
    if return_inputs:
        return out, torch.stack(ins, dim=1)
    return out


def default_args(**kw):
    """argparse defaults the hot path reads (main.py:60-64,79)."""
    d = dict(num_resblock=16, discrim_resblocks=4, discrim_channels=128, crop_size=32, RNN_N=10)
    d.update(kw)
    return types.SimpleNamespace(**d)

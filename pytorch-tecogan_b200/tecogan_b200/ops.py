"""Host-side mirror of the reference ``code/ops.py`` for the hot path.

Same names, argument meaning and error behaviour as dwight-foster/Pytorch-TecoGAN
``code/ops.py`` (cited per function), so ``from ops import *`` keeps working (main.py:28,
code/dataloader.py:1).  Tensor helpers run hand-written sm_100a kernels through the C ABI
(include/tecogan_b200.h); layer factories return the same ``torch.nn`` modules as the
reference so that ``state_dict`` keys, shapes and default initialisation are identical.

New helpers (named by the task's north_star, inline code in the reference):
``warp``, ``space_to_depth``, ``depth_to_space``, ``fused_frame_input``.
"""
import numpy as np  # noqa: F401  (re-exported through the reference's star-import chain)
import torch
import torch.nn as nn
import torch.nn.functional as F  # noqa: F401

from . import _native as _nt


# ----------------------------------------------------------------- preprocessing (ops.py:24-41)
def preprocess(image):
    """[0,1] -> [-1,1]   (code/ops.py:24-26)"""
    return image * 2 - 1


def deprocess(image):
    """[-1,1] -> [0,1]   (code/ops.py:29-31)"""
    return (image + 1) / 2


def preprocessLr(image):
    """identity          (code/ops.py:34-36)"""
    return image


def deprocessLr(image):
    """identity          (code/ops.py:39-41)"""
    return image


# ------------------------------------------------------------------ layer factories (ops.py:45-88)
def conv2_tran(input_channels, kernel=3, output_channel=64, stride=1, use_bias=True, output_padding=0):
    """code/ops.py:45-54"""
    return nn.ConvTranspose2d(input_channels, output_channel, kernel, stride, padding=int((kernel - 1) / 2),
                              bias=bool(use_bias), output_padding=output_padding)


def conv2(batch_input, kernel=3, output_channels=64, stride=1, use_bias=True):
    """code/ops.py:57-63"""
    return nn.Conv2d(batch_input, output_channels, kernel, stride, padding=int((kernel - 1) / 2),
                     bias=bool(use_bias))


def prelu(inputs):
    """code/ops.py:66-68"""
    return nn.PReLU(inputs.shape[1], 0)


def lrelu(alphas):
    """code/ops.py:71-72"""
    return nn.LeakyReLU(negative_slope=alphas)


def batchnorm(inputs, is_training):
    """code/ops.py:75-77 (is_training is ignored by the reference as well)"""
    return nn.BatchNorm2d(inputs, eps=0.001)


def maxpool(kernel_size=(2, 2)):
    """code/ops.py:80-82"""
    return nn.MaxPool2d(kernel_size)


def denselayer(inputs, output_size):
    """code/ops.py:85-88"""
    fc = nn.Linear(inputs, output_size)
    torch.nn.init.xavier_uniform_(fc.weight)
    return fc


# --------------------------------------------------------------------------- tensor helpers
def _to_dev(t):
    """The reference calls some helpers on CPU tensors (main.py:186).  Compute always happens on
    the GPU; the result is returned on the caller's device."""
    return (t, None) if t.is_cuda else (t.cuda(), t.device)


def upscale_four(inputs):
    """nn.Upsample(scale_factor=4, mode="bilinear")   (code/ops.py:98-100)"""
    x, back = _to_dev(inputs)
    x = _nt.require_cuda_f32(x, "upscale_four")
    n, c, h, w = x.shape
    out = torch.empty((n, c, 4 * h, 4 * w), dtype=torch.float32, device=x.device)
    if out.numel():
        _nt.check(_nt.lib().tg_upscale4_bilinear(_nt.ptr(x), _nt.ptr(out), n, c, h, w, 1.0, _nt.stream_ptr()))
    out = out.to(inputs.dtype)
    return out if back is None else out.to(back)


def bicubic_four(inputs):
    """code/ops.py:103-105 — unused by the hot path; kept importable."""
    return nn.Upsample(scale_factor=4, mode="bicubic")(inputs)


def warp(img, grid):
    """F.grid_sample(img, grid.half()) — bilinear, zeros padding, align_corners=False
    (inline in the reference: main.py:203; code/train.py:81,98,165,187).  The grid is rounded
    to fp16 inside the kernel exactly as the reference's ``.half()`` does."""
    x, back = _to_dev(img)
    x = _nt.require_cuda_f32(x, "warp(img)")
    g = _nt.require_cuda_f32(grid.to(x.device), "warp(grid)")
    n, c, h, w = x.shape
    if g.dim() != 4 or g.shape[0] != n or g.shape[3] != 2:
        raise RuntimeError(f"warp: grid must be [N,Ho,Wo,2] with N={n}, got {tuple(g.shape)}")
    ho, wo = g.shape[1], g.shape[2]
    out = torch.empty((n, c, ho, wo), dtype=torch.float32, device=x.device)
    if out.numel():
        _nt.check(_nt.lib().tg_warp_bilinear(_nt.ptr(x), _nt.ptr(g), _nt.ptr(out), n, c, h, w, ho, wo,
                                             _nt.stream_ptr()))
    return out if back is None else out.to(back)


def _as_words(x, name):
    if x.element_size() != 4:
        raise RuntimeError(f"{name}: only 4-byte element types are supported (got {x.dtype})")
    return x.contiguous()


def space_to_depth(x, r=4):
    """x.view(N,C,H,r,W,r).permute(0,1,3,5,2,4).reshape(N,C*r*r,H,W)  (main.py:207-212;
    code/train.py:102-106) == F.pixel_unshuffle(x, r).  Bit-exact data movement."""
    xd, back = _to_dev(x)
    if not xd.is_cuda:
        raise RuntimeError("space_to_depth: CUDA device required")
    xd = _as_words(xd, "space_to_depth")
    n, c, hh, ww = xd.shape
    if hh % r or ww % r:
        raise RuntimeError(f"space_to_depth: spatial size {hh}x{ww} not divisible by {r}")
    out = torch.empty((n, c * r * r, hh // r, ww // r), dtype=xd.dtype, device=xd.device)
    if out.numel():
        _nt.check(_nt.lib().tg_space_to_depth(_nt.ptr(xd), _nt.ptr(out), n, c, hh // r, ww // r, r, _nt.stream_ptr()))
    return out if back is None else out.to(back)


def depth_to_space(x, r=4):
    """inverse of space_to_depth == F.pixel_shuffle(x, r).  Bit-exact data movement."""
    xd, back = _to_dev(x)
    xd = _as_words(xd, "depth_to_space")
    n, crr, h, w = xd.shape
    if crr % (r * r):
        raise RuntimeError(f"depth_to_space: channels {crr} not divisible by {r * r}")
    c = crr // (r * r)
    out = torch.empty((n, c, h * r, w * r), dtype=xd.dtype, device=xd.device)
    if out.numel():
        _nt.check(_nt.lib().tg_depth_to_space(_nt.ptr(xd), _nt.ptr(out), n, c, h, w, r, _nt.stream_ptr()))
    return out if back is None else out.to(back)


def fused_frame_input(lr_t, lr_prev=None, prev_hr=None):
    """Generator input of one frame as NHWC bf16 [N,H,W,64] (main.py:186-213 fused):
    cat(lr_t, space_to_depth(deprocess(warp(prev_hr, flow(lr_prev))))) with the flow computed on
    the fly; lr_prev/prev_hr None -> first frame (zeros, main.py:191-193)."""
    lr_t = _nt.require_cuda_f32(lr_t, "fused_frame_input(lr_t)")
    n, c, h, w = lr_t.shape
    if c != 3:
        raise RuntimeError("fused_frame_input: LR frames must have 3 channels")
    if (lr_prev is None) != (prev_hr is None):
        raise RuntimeError("fused_frame_input: lr_prev and prev_hr go together")
    if lr_prev is not None:
        lr_prev = _nt.require_cuda_f32(lr_prev, "fused_frame_input(lr_prev)")
        prev_hr = _nt.require_cuda_f32(prev_hr, "fused_frame_input(prev_hr)")
        if tuple(prev_hr.shape) != (n, 3, 4 * h, 4 * w) or tuple(lr_prev.shape) != (n, 3, h, w):
            raise RuntimeError("fused_frame_input: shape mismatch")
    x = torch.empty((n, h, w, 64), dtype=torch.bfloat16, device=lr_t.device)
    _nt.check(_nt.lib().tg_fused_warp_s2d_concat(_nt.ptr(lr_t), _nt.ptr(lr_prev), _nt.ptr(prev_hr), _nt.ptr(x), n, h, w,
                                                 3 * h * w, 48 * h * w, _nt.stream_ptr()))
    return x


# ------------------------------------------------------------------ I/O helpers (ops.py:130-242)
def compute_psnr(ref, target):
    """code/ops.py:130-139 (PSNR on the 255 range; never called by the reference)."""
    mse = torch.mean((target.float() - ref.float()) ** 2)
    return 10.0 * torch.log10(255.0 * 255.0 / mse)


def load_ckpt(checkpoint, model):
    """code/ops.py:228-229 — `checkpoint` is a path."""
    return model.load_state_dict(torch.load(checkpoint))


def save_as_gif(tensor, filepath):
    """code/ops.py:234-237 — output helper, outside the hot path; needs imageio."""
    import imageio  # deferred: not needed by the compute path
    images = tensor.clone().detach().cpu().numpy()
    images = (np.transpose(images, (0, 2, 3, 1)) * 255).astype(np.uint8)
    imageio.mimsave(filepath, list(images))


def save_img(out_path, img):
    """code/ops.py:240-242 — output helper, outside the hot path; needs cv2."""
    import cv2
    img = np.clip(img * 255.0, 0, 255).astype(np.uint8)
    cv2.imwrite(out_path, img[:, :, ::-1])

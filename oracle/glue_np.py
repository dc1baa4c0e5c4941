"""numpy restatement of the memory-bound glue (TEST INFRASTRUCTURE, see oracle/__init__.py).

The reference has no helpers for these; it writes them inline and delegates the arithmetic
to PyTorch (un-vendored third-party dependency; README.md:3 pins "pytorch 1.7.1", torch
2.11.0 is what is installed here — the op semantics used are unchanged since 1.3).  Each
function restates the published ATen algorithm for the exact call the reference makes and
cites the reference call site it pins.  All float32, vectorised numpy.
"""
import numpy as np

F32 = np.float32


def space_to_depth(x, r=4):
    """/root/reference/main.py:207-212 and code/train.py:102-106:
        x.view(N,3,H,4,W,4).permute(0,1,3,5,2,4).reshape(N,48,H,W)
    i.e. out[n, c*r*r + dy*r + dx, y, x] = in[n, c, r*y+dy, r*x+dx]  (== F.pixel_unshuffle)."""
    n, c, hh, ww = x.shape
    h, w = hh // r, ww // r
    return np.ascontiguousarray(
        x.reshape(n, c, h, r, w, r).transpose(0, 1, 3, 5, 2, 4)).reshape(n, c * r * r, h, w)


def depth_to_space(x, r=4):
    """Inverse of space_to_depth (== F.pixel_shuffle); not used by the reference, named by
    BASELINE.json north_star; pinned by round trip with space_to_depth."""
    n, crr, h, w = x.shape
    c = crr // (r * r)
    return np.ascontiguousarray(
        x.reshape(n, c, r, r, h, w).transpose(0, 1, 4, 2, 5, 3)).reshape(n, c, h * r, w * r)


def deprocess(x):
    """/root/reference/code/ops.py:29-31  (image + 1) / 2."""
    return ((x.astype(F32) + F32(1.0)) / F32(2.0)).astype(F32)


def preprocess(x):
    """/root/reference/code/ops.py:24-26  image * 2 - 1."""
    return (x.astype(F32) * F32(2.0) - F32(1.0)).astype(F32)


def upscale_four(x):
    """/root/reference/code/ops.py:98-100: nn.Upsample(scale_factor=4, mode='bilinear'),
    align_corners=False.  ATen upsample_bilinear2d: src = (dst+0.5)*0.25-0.5 clamped at 0,
    i0=floor(src), i1=min(i0+1,in-1), l1=src-i0, l0=1-l1,
    out = h0*(w0*p00 + w1*p01) + h1*(w0*p10 + w1*p11)."""
    x = x.astype(F32)
    n, c, h, w = x.shape

    def axis(o, size):
        src = (np.arange(o, dtype=F32) + F32(0.5)) * F32(0.25) - F32(0.5)
        src = np.maximum(src, F32(0.0)).astype(F32)
        i0 = np.floor(src).astype(np.int64)
        i0 = np.minimum(i0, size - 1)
        i1 = np.minimum(i0 + 1, size - 1)
        l1 = (src - i0.astype(F32)).astype(F32)
        l0 = (F32(1.0) - l1).astype(F32)
        return i0, i1, l0, l1

    y0, y1, hl0, hl1 = axis(4 * h, h)
    x0, x1, wl0, wl1 = axis(4 * w, w)
    p00 = x[:, :, y0][:, :, :, x0]
    p01 = x[:, :, y0][:, :, :, x1]
    p10 = x[:, :, y1][:, :, :, x0]
    p11 = x[:, :, y1][:, :, :, x1]
    wl0 = wl0[None, None, None, :]
    wl1 = wl1[None, None, None, :]
    hl0 = hl0[None, None, :, None]
    hl1 = hl1[None, None, :, None]
    top = (wl0 * p00 + wl1 * p01).astype(F32)
    bot = (wl0 * p10 + wl1 * p11).astype(F32)
    return (hl0 * top + hl1 * bot).astype(F32)


def warp(img, grid):
    """/root/reference/main.py:203, code/train.py:98,165,187:
        F.grid_sample(img, grid.half())   (bilinear, padding_mode='zeros', align_corners=False)
    The grid is rounded to fp16 first (that rounding is part of the reference numerics), then
    ATen grid_sampler_2d: ix = ((gx+1)*W-1)/2, 4 taps, each masked by its own bounds test."""
    img = img.astype(F32)
    g = grid.astype(np.float16).astype(F32)
    n, c, h, w = img.shape
    gx, gy = g[..., 0], g[..., 1]
    ix = (((gx + F32(1.0)) * F32(w) - F32(1.0)) / F32(2.0)).astype(F32)
    iy = (((gy + F32(1.0)) * F32(h) - F32(1.0)) / F32(2.0)).astype(F32)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    x1 = x0 + 1
    y1 = y0 + 1
    w_nw = ((x1 - ix) * (y1 - iy)).astype(F32)
    w_ne = ((ix - x0) * (y1 - iy)).astype(F32)
    w_sw = ((x1 - ix) * (iy - y0)).astype(F32)
    w_se = ((ix - x0) * (iy - y0)).astype(F32)
    out = np.zeros((n, c) + gx.shape[1:], dtype=F32)
    nn = np.arange(n)[:, None, None]

    def tap(xx, yy, wt):
        ok = (xx >= 0) & (xx <= w - 1) & (yy >= 0) & (yy <= h - 1)
        xi = np.clip(xx, 0, w - 1).astype(np.int64)
        yi = np.clip(yy, 0, h - 1).astype(np.int64)
        v = img[nn, :, yi, xi]                      # [n, Ho, Wo, c]
        v = np.where(ok[..., None], v, F32(0.0))
        return (v * wt[..., None]).astype(F32)

    acc = tap(x0, y0, w_nw)
    acc = (acc + tap(x1, y0, w_ne)).astype(F32)
    acc = (acc + tap(x0, y1, w_sw)).astype(F32)
    acc = (acc + tap(x1, y1, w_se)).astype(F32)
    out[:] = acc.transpose(0, 3, 1, 2)
    return out


def flow_from_lr(lr_prev):
    """/root/reference/main.py:186-189,200-201: the 'flow' of frame i is
    upscale_four(LR_i*4)[:,0:2] whose [2,Ho,Wo] memory is re-interpreted (``.view``, not
    permute) as a [Ho,Wo,2] sampling grid."""
    n, _, h, w = lr_prev.shape
    up = upscale_four(lr_prev.astype(F32) * F32(4.0))[:, 0:2]
    return np.ascontiguousarray(up).reshape(n, 4 * h, 4 * w, 2)


def frame_input(lr_t, lr_prev, prev_hr):
    """/root/reference/main.py:199-213 — generator input for frame t>0:
    cat(LR_t, space_to_depth(deprocess(warp(HR_{t-1}, flow_{t-1}))))  -> [N,51,H,W]."""
    wp = warp(prev_hr, flow_from_lr(lr_prev))
    return np.concatenate([lr_t.astype(F32), space_to_depth(deprocess(wp), 4)], axis=1)

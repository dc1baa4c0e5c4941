"""GPU parity of the tcgen05 convolution kernels (through the C ABI) against torch CPU fp32
convolutions of the same bf16-rounded operands.  Tolerance: bf16 conv path, <=1e-2 relative
max-abs (BASELINE.json north_star); here operands are pre-rounded so only fp32 accumulation
order and the bf16 output rounding differ -> much tighter bound used."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu

HALO, DX3 = 0, 1


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc_bf16(x_nchw, cpad):
    n, c, h, w = x_nchw.shape
    out = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.cuda()


def _pack(kind, w, b, cin, cout):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    packed = torch.zeros(lib.tg_packed_conv_bytes(kind, cin, cout), dtype=torch.uint8, device="cuda")
    wd = w.contiguous().cuda()
    bd = b.contiguous().cuda() if b is not None else None
    nt.check(lib.tg_pack_weights(kind, nt.ptr(wd), nt.ptr(bd), cin, cout, nt.ptr(packed), nt.stream_ptr()))
    torch.cuda.synchronize()
    return packed


def _rel(got, want):
    return (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)


@pytest.mark.parametrize("amode", [HALO, DX3], ids=["halo", "dx3"])
@pytest.mark.parametrize("n,h,w,cin,cout,relu,resid", [
    (1, 16, 8, 64, 64, 1, 0),      # exactly one work item
    (1, 20, 13, 64, 64, 0, 1),     # ragged edges + residual
    (2, 33, 40, 51, 64, 1, 0),     # padded input channels, batch
    (1, 24, 24, 64, 128, 1, 0),    # two output chunks
    (1, 24, 24, 128, 64, 1, 0),    # two K chunks
    (1, 17, 9, 128, 128, 0, 0),
    (1, 180, 320, 64, 64, 0, 1),   # cfg2 trunk layer, persistent CTAs loop over many items
])
def test_conv3x3(amode, n, h, w, cin, cout, relu, resid):
    from tecogan_b200 import _native as nt
    if amode == DX3 and cin > 64:
        pytest.skip("DX3 staging does not fit two K chunks (diagnostic mode only)")
    lib = nt.lib()
    cpad = 64 if cin <= 64 else 128
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 1, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 3, 3), 2, -0.1, 0.1)))
    b = torch.from_numpy(synth.det_uniform((cout,), 3, -0.5, 0.5))
    r = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 4, -1, 1))) if resid else None
    want = F.conv2d(x, wt, b, padding=1)
    if relu:
        want = want.relu()
    if resid:
        want = want + r
    packed = _pack(0, wt, b, cin, cout)
    xd = _nhwc_bf16(x, cpad)
    rd = _nhwc_bf16(r, cout) if resid else None
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_conv3x3_fwd(nt.ptr(xd), nt.ptr(packed), nt.ptr(rd), nt.ptr(y), n, h, w, cpad, cout, relu, amode,
                                nt.stream_ptr()))
    torch.cuda.synchronize()
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert _rel(got, want) <= 6e-3, _rel(got, want)      # bf16 output rounding: 2^-9 relative


@pytest.mark.parametrize("amode", [HALO, DX3], ids=["halo", "dx3"])
@pytest.mark.parametrize("n,h,w,c", [(1, 16, 8, 64), (2, 19, 21, 64), (1, 12, 20, 128), (1, 90, 160, 64)])
def test_conv_transpose(amode, n, h, w, c):
    from tecogan_b200 import _native as nt
    if amode == DX3 and c > 64:
        pytest.skip("DX3 staging does not fit two K chunks (diagnostic mode only)")
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, c, h, w), 1, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((c, c, 3, 3), 2, -0.1, 0.1)))      # [cin,cout,3,3]
    b = torch.from_numpy(synth.det_uniform((c,), 3, -0.5, 0.5))
    want = F.conv_transpose2d(x, wt, b, stride=2, padding=1, output_padding=1).relu()
    packed = _pack(1, wt, b, c, c)
    xd = _nhwc_bf16(x, c)
    y = torch.empty(n, 2 * h, 2 * w, c, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_convT3x3s2_fwd(nt.ptr(xd), nt.ptr(packed), nt.ptr(y), n, h, w, c, c, 1, amode, nt.stream_ptr()))
    torch.cuda.synchronize()
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert _rel(got, want) <= 6e-3, _rel(got, want)


@pytest.mark.parametrize("amode", [HALO, DX3], ids=["halo", "dx3"])
@pytest.mark.parametrize("n,h,w", [(1, 16, 8), (2, 37, 50), (1, 256, 256)])
def test_output_conv_sigmoid(amode, n, h, w):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, 64, h, w), 1, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((3, 64, 3, 3), 2, -0.2, 0.2)))
    b = torch.from_numpy(synth.det_uniform((3,), 3, -0.5, 0.5))
    logits = F.conv2d(x, wt, b, padding=1)
    packed = _pack(0, wt, b, 64, 3)
    xd = _nhwc_bf16(x, 64)
    out = torch.empty(n, 3, h, w, dtype=torch.float32, device="cuda")
    lg = torch.empty_like(out)
    nt.check(lib.tg_conv3x3_out_sigmoid(nt.ptr(xd), nt.ptr(packed), nt.ptr(out), nt.ptr(lg), n, h, w, amode,
                                        nt.stream_ptr()))
    torch.cuda.synchronize()
    assert (lg.cpu() - logits).abs().max().item() <= 1e-4          # fp32 accumulate, fp32 store
    assert (out.cpu() - torch.sigmoid(logits)).abs().max().item() <= 1e-5


def test_conv_full_size_windows():
    # cfg2 4x-resolution layer size (128->64 at 720x1280): too big for a full CPU oracle pass in a test, so
    # sampled 24x24 windows (corners, interior, edges) are checked against torch CPU on the same operands
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    n, h, w, cin, cout = 1, 720, 1280, 128, 64
    torch.manual_seed(0)
    a = (torch.rand(n, h, w, cin, device="cuda") - 0.5).to(torch.bfloat16)
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 3, 3), 2, -0.05, 0.05)))
    packed = _pack(0, wt, None, cin, cout)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_conv3x3_fwd(nt.ptr(a), nt.ptr(packed), None, nt.ptr(y), n, h, w, cin, cout, 0, HALO, nt.stream_ptr()))
    torch.cuda.synchronize()
    for (y0, x0) in [(0, 0), (352, 640), (720 - 24, 1280 - 24), (100, 1280 - 24)]:
        win = a[:, max(y0 - 1, 0):y0 + 25, max(x0 - 1, 0):x0 + 25].float().cpu().permute(0, 3, 1, 2)
        ref = F.conv2d(win, wt, None, padding=1)
        oy, ox = (1 if y0 > 0 else 0), (1 if x0 > 0 else 0)
        ref = ref[:, :, oy:oy + 24, ox:ox + 24]
        got = y[:, y0:y0 + 24, x0:x0 + 24].float().cpu().permute(0, 3, 1, 2)
        assert _rel(got, ref) <= 6e-3


@pytest.mark.parametrize("n,h,w,cin,cout,act", [
    (1, 32, 16, 64, 64, 0),        # exactly one output tile
    (2, 38, 26, 64, 64, 2),        # ragged edges, LeakyReLU(0.2)
    (1, 64, 64, 64, 128, 0),       # block2: two output chunks
    (3, 32, 32, 128, 128, 0),      # block3: two K chunks
    (2, 16, 16, 128, 64, 1),       # block4
    (12, 128, 128, 64, 64, 0),     # block1 at the cfg4 size
])
def test_conv4x4s2(n, h, w, cin, cout, act):
    """Conv2d(k=4, s=2, p=1, bias=False) of discriminator_block (code/models.py:90-94): strided TMA phase boxes."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 1, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 4, 4), 2, -0.1, 0.1)))
    want = F.conv2d(x, wt, None, stride=2, padding=1)
    if act == 1:
        want = want.relu()
    elif act == 2:
        want = F.leaky_relu(want, 0.2)
    packed = _pack(2, wt, None, cin, cout)
    xd = _nhwc_bf16(x, cin)
    y = torch.empty(n, h // 2, w // 2, cout, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_conv4x4s2_fwd(nt.ptr(xd), nt.ptr(packed), nt.ptr(y), n, h, w, cin, cout, act, nt.stream_ptr()))
    torch.cuda.synchronize()
    got = y.float().cpu().permute(0, 3, 1, 2)
    assert got.shape == want.shape
    assert _rel(got, want) <= 6e-3, _rel(got, want)


@pytest.mark.parametrize("n,h,w", [(3, 8, 8), (2, 16, 16), (1, 34, 18)])
def test_conv4x4s2_three_channels_raw_nchw(n, h, w):
    """discriminator block5 (64 -> 3): raw f32 NCHW output feeding BatchNorm + flatten (code/models.py:121,141-142)."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, 64, h, w), 5, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((3, 64, 4, 4), 6, -0.1, 0.1)))
    want = F.conv2d(x, wt, None, stride=2, padding=1)
    packed = _pack(2, wt, None, 64, 3)
    xd = _nhwc_bf16(x, 64)
    y = torch.empty(n, 3, h // 2, w // 2, dtype=torch.float32, device="cuda")
    nt.check(lib.tg_conv4x4s2_fwd(nt.ptr(xd), nt.ptr(packed), nt.ptr(y), n, h, w, 64, 3, 0, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert (y.cpu() - want).abs().max().item() <= 1e-4


def test_conv3x3_leaky_relu():
    """discriminator first conv: conv3x3(27 -> 64) + bias + LeakyReLU(0.2) (code/models.py:102)."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    n, h, w, cin, cout = 2, 24, 40, 27, 64
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 1, -1, 1)))
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 3, 3), 2, -0.1, 0.1)))
    b = torch.from_numpy(synth.det_uniform((cout,), 3, -0.5, 0.5))
    want = F.leaky_relu(F.conv2d(x, wt, b, padding=1), 0.2)
    packed = _pack(0, wt, b, cin, cout)
    xd = _nhwc_bf16(x, 64)
    y = torch.empty(n, h, w, cout, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_conv3x3_fwd(nt.ptr(xd), nt.ptr(packed), None, nt.ptr(y), n, h, w, 64, cout, 2, HALO, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(y.float().cpu().permute(0, 3, 1, 2), want) <= 6e-3

"""GPU parity of the backward (training) kernels against torch CPU autograd on the same bf16-rounded operands."""
import pytest
import torch
import torch.nn.functional as F

from oracle import synth

pytestmark = pytest.mark.gpu


def _bf(x):
    return x.to(torch.bfloat16).float()


def _nhwc_bf16(x_nchw, cpad):
    n, c, h, w = x_nchw.shape
    out = torch.zeros(n, h, w, cpad, dtype=torch.bfloat16)
    out[..., :c] = x_nchw.permute(0, 2, 3, 1).to(torch.bfloat16)
    return out.cuda()


def _rel(got, want):
    return (got - want).abs().max().item() / max(want.abs().max().item(), 1e-6)


@pytest.mark.parametrize("n,h,w,cin,cout", [
    (1, 16, 8, 64, 64),        # one tile
    (2, 20, 13, 64, 64),       # ragged edges, batch
    (1, 32, 32, 51, 64),       # padded input channels (generator conv.0)
    (1, 24, 24, 64, 128),      # 128 output channels: two dY boxes
    (1, 24, 24, 128, 64),      # 128 input channels: one MMA per kx tap
    (2, 17, 9, 128, 128),
    (1, 40, 40, 64, 3),        # output conv
    (4, 64, 64, 64, 64),       # many tiles per CTA, several slabs
])
def test_conv3x3_wgrad(n, h, w, cin, cout):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 1, -1, 1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 2, -1, 1)))
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    F.conv2d(x, wt, None, padding=1).backward(dy)
    want = wt.grad
    xd = _nhwc_bf16(x, 64 if cin <= 64 else 128)
    dyd = _nhwc_bf16(dy, 64 if cout <= 64 else 128)
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    nt.check(lib.tg_conv3x3_wgrad(nt.ptr(xd), nt.ptr(dyd), nt.ptr(dw), n, h, w, cin, cout, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), want) <= 1e-4, _rel(dw.cpu(), want)       # exact bf16 products, fp32 accumulation
    # accumulates: a second call doubles the result
    nt.check(lib.tg_conv3x3_wgrad(nt.ptr(xd), nt.ptr(dyd), nt.ptr(dw), n, h, w, cin, cout, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), 2 * want) <= 1e-4

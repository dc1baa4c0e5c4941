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


def _pack(kind, w, cin, cout):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    packed = torch.zeros(lib.tg_packed_conv_bytes(kind, cin, cout), dtype=torch.uint8, device="cuda")
    wd = w.contiguous().cuda()
    nt.check(lib.tg_pack_weights(kind, nt.ptr(wd), None, cin, cout, nt.ptr(packed), nt.stream_ptr()))
    torch.cuda.synchronize()
    return packed


@pytest.mark.parametrize("n,h,w,cin,cout,use_mask,use_resid", [
    (1, 16, 8, 64, 64, 0, 0),
    (2, 20, 13, 64, 64, 1, 0),     # ReLU backward mask
    (1, 24, 24, 64, 128, 0, 1),    # + gradient of the skip connection
    (1, 24, 24, 128, 64, 1, 1),
    (2, 17, 9, 128, 128, 0, 0),
    (1, 40, 40, 64, 3, 1, 0),      # output conv: dY has 3 real channels
])
def test_conv3x3_dgrad(n, h, w, cin, cout, use_mask, use_resid):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = torch.zeros(n, cin, h, w, requires_grad=True)
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 3, 3), 2, -0.1, 0.1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 3, -1, 1)))
    F.conv2d(x, wt, None, padding=1).backward(dy)
    want = x.grad
    m = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 4, -1, 1))).relu() if use_mask else None
    r = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 5, -1, 1))) if use_resid else None
    if use_resid:
        want = want + r
    if use_mask:
        want = want * (m > 0)
    packed = _pack(3, wt, cin, cout)
    dyd = _nhwc_bf16(dy, 64 if cout <= 64 else 128)
    dx = torch.empty(n, h, w, cin, dtype=torch.bfloat16, device="cuda")
    rd = _nhwc_bf16(r, cin) if use_resid else None          # keep the device tensors alive across the launch
    md = _nhwc_bf16(m, cin) if use_mask else None
    nt.check(lib.tg_conv3x3_dgrad(nt.ptr(dyd), nt.ptr(packed), nt.ptr(rd), nt.ptr(md), nt.ptr(dx), n, h, w, cin, cout,
                                  nt.stream_ptr()))
    torch.cuda.synchronize()
    got = dx.float().cpu().permute(0, 3, 1, 2)
    assert _rel(got, want) <= 6e-3, _rel(got, want)


@pytest.mark.parametrize("n,h,w,c,use_mask", [(1, 16, 8, 64, 0), (2, 19, 21, 64, 1), (1, 12, 20, 128, 1)])
def test_conv_transpose_dgrad(n, h, w, c, use_mask):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = torch.zeros(n, c, h, w, requires_grad=True)
    wt = _bf(torch.from_numpy(synth.det_uniform((c, c, 3, 3), 2, -0.1, 0.1)))          # [cin,cout,3,3]
    dy = _bf(torch.from_numpy(synth.det_uniform((n, c, 2 * h, 2 * w), 3, -1, 1)))
    F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1).backward(dy)
    want = x.grad
    m = _bf(torch.from_numpy(synth.det_uniform((n, c, h, w), 4, -1, 1))).relu() if use_mask else None
    if use_mask:
        want = want * (m > 0)
    packed = _pack(4, wt, c, c)
    dyd = _nhwc_bf16(dy, c)
    dx = torch.empty(n, h, w, c, dtype=torch.bfloat16, device="cuda")
    md = _nhwc_bf16(m, c) if use_mask else None
    nt.check(lib.tg_convT3x3s2_dgrad(nt.ptr(dyd), nt.ptr(packed), nt.ptr(md), nt.ptr(dx), n, h, w, c, c, nt.stream_ptr()))
    torch.cuda.synchronize()
    got = dx.float().cpu().permute(0, 3, 1, 2)
    assert _rel(got, want) <= 6e-3, _rel(got, want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 8, 64, 64), (2, 19, 21, 64, 64), (1, 12, 20, 128, 128), (2, 32, 32, 64, 64)])
def test_conv_transpose_wgrad(n, h, w, cin, cout):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 1, -1, 1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, 2 * h, 2 * w), 2, -1, 1)))
    wt = torch.zeros(cin, cout, 3, 3, requires_grad=True)
    F.conv_transpose2d(x, wt, None, stride=2, padding=1, output_padding=1).backward(dy)
    want = wt.grad
    xd, dyd = _nhwc_bf16(x, cin), _nhwc_bf16(dy, cout)
    dw = torch.zeros(cin, cout, 3, 3, device="cuda")
    nt.check(lib.tg_convT3x3s2_wgrad(nt.ptr(xd), nt.ptr(dyd), nt.ptr(dw), n, h, w, cin, cout, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), want) <= 1e-4, _rel(dw.cpu(), want)


@pytest.mark.parametrize("n,h,w,cin,cout", [(1, 16, 8, 64, 64), (2, 19, 13, 64, 128), (3, 16, 16, 128, 128), (2, 8, 8, 128, 64),
                                            (2, 4, 4, 64, 3), (12, 64, 64, 64, 64)])
def test_conv4x4s2_wgrad(n, h, w, cin, cout):
    """h, w = output size of the stride-2 conv; x is [n,cin,2h,2w]."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, 2 * h, 2 * w), 1, -1, 1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 2, -1, 1)))
    wt = torch.zeros(cout, cin, 4, 4, requires_grad=True)
    F.conv2d(x, wt, None, stride=2, padding=1).backward(dy)
    want = wt.grad
    xd, dyd = _nhwc_bf16(x, cin), _nhwc_bf16(dy, 64 if cout <= 64 else 128)
    dw = torch.zeros(cout, cin, 4, 4, device="cuda")
    nt.check(lib.tg_conv4x4s2_wgrad(nt.ptr(xd), nt.ptr(dyd), nt.ptr(dw), n, h, w, cin, cout, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), want) <= 1e-4, _rel(dw.cpu(), want)


@pytest.mark.parametrize("pixels,c", [(100, 64), (4096, 128), (777, 3)])
def test_bias_grad(pixels, c):
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    cpad = 64 if c <= 64 else 128
    dy = torch.zeros(pixels, cpad)
    dy[:, :c] = _bf(torch.from_numpy(synth.det_uniform((pixels, c), 7, -1, 1)))
    db = torch.zeros(c, device="cuda")
    dyd = dy.to(torch.bfloat16).cuda()
    nt.check(lib.tg_bias_grad(nt.ptr(dyd), nt.ptr(db), pixels, c, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert (db.cpu() - dy[:, :c].sum(0)).abs().max().item() <= 1e-3 * max(1.0, pixels ** 0.5)


@pytest.mark.parametrize("n,h,w,cin,cout,mask_mode", [(1, 16, 8, 64, 64, 0), (2, 19, 13, 64, 128, 2), (3, 16, 16, 128, 128, 1),
                                                      (2, 8, 8, 128, 64, 0), (2, 4, 4, 64, 3, 2)])
def test_conv4x4s2_dgrad(n, h, w, cin, cout, mask_mode):
    """h, w = output size of the forward stride-2 conv; dX is [n,cin,2h,2w]."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = torch.zeros(n, cin, 2 * h, 2 * w, requires_grad=True)
    wt = _bf(torch.from_numpy(synth.det_uniform((cout, cin, 4, 4), 2, -0.1, 0.1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 3, -1, 1)))
    F.conv2d(x, wt, None, stride=2, padding=1).backward(dy)
    want = x.grad
    m = _bf(torch.from_numpy(synth.det_uniform((n, cin, 2 * h, 2 * w), 4, -1, 1)))
    if mask_mode == 1:
        m = m.relu()
        want = want * (m > 0)
    elif mask_mode == 2:
        m = F.leaky_relu(m, 0.2)
        want = want * torch.where(m > 0, 1.0, 0.2)
    packed = _pack(5, wt, cin, cout)
    dyd = _nhwc_bf16(dy, 64 if cout <= 64 else 128)
    md = _nhwc_bf16(m, cin) if mask_mode else None
    dx = torch.empty(n, 2 * h, 2 * w, cin, dtype=torch.bfloat16, device="cuda")
    nt.check(lib.tg_conv4x4s2_dgrad(nt.ptr(dyd), nt.ptr(packed), nt.ptr(md), mask_mode, nt.ptr(dx), n, h, w, cin, cout,
                                    nt.stream_ptr()))
    torch.cuda.synchronize()
    got = dx.float().cpu().permute(0, 3, 1, 2)
    assert _rel(got, want) <= 6e-3, _rel(got, want)


@pytest.mark.parametrize("pixels,c,act,skip", [(4096, 64, 2, False), (1000, 128, 0, True), (12 * 64 * 64, 64, 2, False),
                                               (37, 128, 0, False)])
def test_bn_abi_stats_apply_bwd_vs_torch(pixels, c, act, skip):
    """tg_bn_stats / tg_bn_apply / tg_bn_bwd (nn.BatchNorm2d(eps=1e-3) in training mode + LeakyReLU(0.2) /
    skip, code/ops.py:75-77, code/models.py:90-94,106,130) against torch fp32 autograd on the same NHWC tensors."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = torch.from_numpy(synth.det_uniform((pixels, c), 5, -2.0, 3.0))
    sk = torch.from_numpy(synth.det_uniform((pixels, c), 6, -1.0, 1.0)) if skip else None
    gamma = torch.from_numpy(synth.det_uniform((c,), 7, 0.5, 1.5))
    beta = torch.from_numpy(synth.det_uniform((c,), 8, -0.3, 0.3))
    g_out = _bf(torch.from_numpy(synth.det_uniform((pixels, c), 9, -1.0, 1.0)))
    bn = torch.nn.BatchNorm2d(c, eps=1e-3)
    with torch.no_grad():
        bn.weight.copy_(gamma)
        bn.bias.copy_(beta)
    bn.train()
    xr = x.clone().requires_grad_(True)
    y = bn(xr.t().reshape(1, c, pixels, 1))
    if act == 2:
        y = F.leaky_relu(y, 0.2)
    y = y.reshape(c, pixels).t()
    if skip:
        y = y + sk
    y.backward(g_out)
    ws = torch.zeros(lib.tg_workspace_bytes_bn(), dtype=torch.uint8, device="cuda")
    stats = torch.empty((c, 4), device="cuda")
    rm, rv = torch.zeros(c, device="cuda"), torch.ones(c, device="cuda")
    nbt = torch.zeros((), dtype=torch.int64, device="cuda")
    xd, gd, bd = x.cuda(), gamma.cuda(), beta.cuda()
    nt.check(lib.tg_bn_stats(nt.ptr(xd), pixels, c, nt.ptr(gd), nt.ptr(bd), nt.ptr(stats), nt.ptr(rm), nt.ptr(rv), nt.ptr(nbt),
                             nt.ptr(ws), ws.numel(), nt.stream_ptr()))
    y32 = torch.empty((pixels, c), device="cuda")
    y16 = torch.empty((pixels, c), dtype=torch.bfloat16, device="cuda")
    skd = sk.cuda() if skip else None
    nt.check(lib.tg_bn_apply(nt.ptr(xd), nt.ptr(skd), nt.ptr(y32), nt.ptr(y16), pixels, c, nt.ptr(stats), act, nt.stream_ptr()))
    assert _rel(y32.cpu(), y.detach()) <= 2e-5
    assert torch.equal(y16.cpu(), y32.cpu().to(torch.bfloat16))
    assert int(nbt) == 1
    assert (rm.cpu() - bn.running_mean).abs().max().item() <= 1e-5 and _rel(rv.cpu(), bn.running_var) <= 1e-4
    # only one of the two outputs requested
    y_only = torch.empty((pixels, c), device="cuda")
    nt.check(lib.tg_bn_apply(nt.ptr(xd), nt.ptr(skd), nt.ptr(y_only), None, pixels, c, nt.ptr(stats), act, nt.stream_ptr()))
    assert torch.equal(y_only, y32)
    dx = torch.empty((pixels, c), dtype=torch.bfloat16, device="cuda")
    dg, db = torch.zeros(c, device="cuda"), torch.zeros(c, device="cuda")
    act_out = y32 if act == 2 else None                          # the block's f32 output (LeakyReLU sign test)
    nt.check(lib.tg_bn_bwd(nt.ptr(g_out.to(torch.bfloat16).cuda()), nt.ptr(xd), nt.ptr(act_out), nt.ptr(dx), pixels, c,
                           nt.ptr(stats), nt.ptr(dg), nt.ptr(db), nt.ptr(ws), ws.numel(), nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dg.cpu(), bn.weight.grad) <= 2e-3 and _rel(db.cpu(), bn.bias.grad) <= 2e-3
    assert _rel(dx.float().cpu(), xr.grad) <= 1.5e-2            # bf16 output
    # workspace too small -> TG_ERR_WORKSPACE, not a crash
    rc = lib.tg_bn_stats(nt.ptr(xd), pixels, c, nt.ptr(gd), nt.ptr(bd), nt.ptr(stats), None, None, None, nt.ptr(ws), 16, nt.stream_ptr())
    assert rc == -4


@pytest.mark.parametrize("n,h,w,cin,cout", [(2, 20, 13, 64, 64), (1, 32, 32, 128, 64), (3, 16, 24, 51, 64), (1, 40, 24, 64, 3)])
def test_conv3x3_wgrad_with_fused_bias_grad_through_generator_backward_shapes(n, h, w, cin, cout):
    """The <= 64-output-channel weight-gradient kernel (filter rows stacked along M) also sums dY into the bias gradient
    while the MMAs run: check both against torch autograd on the same bf16-rounded operands, via the public wgrad entry
    point for dW and tg_bias_grad's contract (db += sum over pixels of dY) for the bias, using the fused path exactly as
    tg_gen_backward / tg_disc_backward drive it (launch through tg_conv3x3_wgrad_bias)."""
    from tecogan_b200 import _native as nt
    lib = nt.lib()
    x = _bf(torch.from_numpy(synth.det_uniform((n, cin, h, w), 11, -1, 1)))
    dy = _bf(torch.from_numpy(synth.det_uniform((n, cout, h, w), 12, -1, 1)))
    wt = torch.zeros(cout, cin, 3, 3, requires_grad=True)
    bt = torch.zeros(cout, requires_grad=True)
    F.conv2d(x, wt, bt, padding=1).backward(dy)
    xd = _nhwc_bf16(x, 64 if cin <= 64 else 128)
    dyd = _nhwc_bf16(dy, 64)
    dw = torch.zeros(cout, cin, 3, 3, device="cuda")
    db = torch.zeros(cout, device="cuda")
    nt.check(lib.tg_conv3x3_wgrad_bias(nt.ptr(xd), nt.ptr(dyd), nt.ptr(dw), nt.ptr(db), n, h, w, cin, cout, nt.stream_ptr()))
    torch.cuda.synchronize()
    assert _rel(dw.cpu(), wt.grad) <= 1e-4, _rel(dw.cpu(), wt.grad)
    assert _rel(db.cpu(), bt.grad) <= 1e-4, _rel(db.cpu(), bt.grad)

"""GPU parity of the glue kernels (through the C ABI) against the CPU oracle.
Bar: bit-exact for space_to_depth / depth_to_space, <=1e-5 max-abs for the fp32 warp,
<=1e-6 for the bilinear x4 upscale (BASELINE.json north_star)."""
import numpy as np
import pytest
import torch

from oracle import glue_np, synth, tecogan_oracle as O

pytestmark = pytest.mark.gpu


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


@pytest.mark.parametrize("shape,r", [((2, 3, 16, 24), 4), ((1, 3, 64, 64), 4), ((1, 1, 4, 4), 4),
                                     ((3, 5, 12, 20), 4), ((2, 3, 6, 10), 2), ((1, 2, 9, 15), 3),
                                     ((1, 3, 720, 1280), 4)])
def test_space_to_depth_bit_exact(shape, r):
    from tecogan_b200 import ops
    x = synth.det_uniform(shape, 5, -1, 1)
    want = glue_np.space_to_depth(x, r)
    got = ops.space_to_depth(_cuda(x), r).cpu().numpy()
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    back = ops.depth_to_space(_cuda(want), r).cpu().numpy()
    assert np.array_equal(back.view(np.uint32), x.view(np.uint32))


def test_space_to_depth_int32_and_empty():
    from tecogan_b200 import ops
    x = torch.arange(2 * 3 * 8 * 8, dtype=torch.int32).reshape(2, 3, 8, 8).cuda()
    want = torch.nn.functional.pixel_unshuffle(x.cpu().float(), 4).to(torch.int32)
    assert torch.equal(ops.space_to_depth(x, 4).cpu(), want)
    e = torch.zeros(0, 3, 8, 8, device="cuda")
    assert ops.space_to_depth(e, 4).shape == (0, 48, 2, 2)
    with pytest.raises(RuntimeError):
        ops.space_to_depth(torch.zeros(1, 3, 7, 8, device="cuda"), 4)


def test_s2d_full_size_round_trip_and_checksum():
    # cfg3 size (4K): size-independent properties — round trip identity and a permutation-invariant checksum
    from tecogan_b200 import ops
    x = torch.rand(1, 3, 2160, 3840, device="cuda")
    y = ops.space_to_depth(x, 4)
    assert y.shape == (1, 48, 540, 960)
    assert torch.equal(ops.depth_to_space(y, 4), x)
    assert torch.equal(y.view(torch.int32).sum(dtype=torch.int64), x.view(torch.int32).sum(dtype=torch.int64))
    assert torch.equal(y, torch.nn.functional.pixel_unshuffle(x, 4))


@pytest.mark.parametrize("n,c,h,w,ho,wo,lim", [(2, 3, 16, 24, 16, 24, 1.2), (1, 3, 33, 17, 20, 44, 1.0),
                                               (1, 1, 8, 8, 5, 3, 3.0), (2, 3, 64, 64, 64, 64, 0.9)])
def test_warp_vs_oracle(n, c, h, w, ho, wo, lim):
    from tecogan_b200 import ops
    img = synth.det_uniform((n, c, h, w), 11, -1, 1)
    grid = synth.det_uniform((n, ho, wo, 2), 12, -lim, lim)
    want = glue_np.warp(img, grid)
    got = ops.warp(_cuda(img), _cuda(grid)).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5
    want_t = O.warp(torch.from_numpy(img), torch.from_numpy(grid)).numpy()
    assert np.abs(got - want_t).max() <= 1e-5


def test_warp_golden(golden_dir):
    import os
    from tecogan_b200 import ops
    g = np.load(os.path.join(golden_dir, "glue.npz"))
    img = synth.det_uniform((2, 3, 16, 24), 11, -1.0, 1.0)
    grid = synth.det_uniform((2, 16, 24, 2), 12, -1.2, 1.2)
    lr = synth.det_uniform((2, 3, 6, 10), 13, 0.0, 1.0)
    assert np.abs(ops.warp(_cuda(img), _cuda(grid)).cpu().numpy() - g["warp"]).max() <= 1e-5
    assert np.abs(ops.upscale_four(_cuda(lr * np.float32(4))).cpu().numpy() - g["upscale"]).max() <= 1e-6
    assert np.array_equal(ops.space_to_depth(_cuda(img), 4).cpu().numpy(), g["s2d"])


def test_warp_identity_and_far_out_of_bounds():
    from tecogan_b200 import ops
    n, c, h, w = 1, 3, 32, 48
    img = torch.rand(n, c, h, w, device="cuda")
    # far outside / non-finite grid -> zeros
    g = torch.full((n, h, w, 2), 7.5, device="cuda")
    assert ops.warp(img, g).abs().max().item() == 0.0
    g[..., 0] = float("inf")
    assert torch.isfinite(ops.warp(img, g)).all()
    # linearity in the image
    grid = (torch.rand(n, h, w, 2, device="cuda") * 2 - 1)
    a, b = torch.rand_like(img), torch.rand_like(img)
    lhs = ops.warp(a + b, grid)
    rhs = ops.warp(a, grid) + ops.warp(b, grid)
    assert (lhs - rhs).abs().max().item() <= 1e-5


@pytest.mark.parametrize("shape", [(2, 3, 6, 10), (1, 3, 1, 1), (1, 2, 7, 13), (4, 3, 32, 32), (1, 3, 180, 320)])
def test_upscale_four_vs_oracle(shape):
    from tecogan_b200 import ops
    x = synth.det_uniform(shape, 13, 0, 4)
    want = glue_np.upscale_four(x)
    got = ops.upscale_four(_cuda(x)).cpu().numpy()
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-6
    want_t = O.upscale_four(torch.from_numpy(x)).numpy()
    assert np.abs(got - want_t).max() <= 2e-6
    # CPU tensor in -> CPU tensor out (reference main.py:186 calls it on CPU tensors)
    assert not ops.upscale_four(torch.from_numpy(x)).is_cuda


@pytest.mark.parametrize("n,h,w,hi", [(1, 16, 16, 0.25), (2, 9, 21, 0.25), (1, 32, 40, 1.0), (3, 8, 8, 0.25)])
def test_fused_frame_input_vs_oracle(n, h, w, hi):
    from tecogan_b200 import ops
    lr_t = synth.det_uniform((n, 3, h, w), 41, 0, hi)
    lr_p = synth.det_uniform((n, 3, h, w), 42, 0, hi)
    hr = synth.det_uniform((n, 3, 4 * h, 4 * w), 43, 0, 1)
    want = O.frame_input(torch.from_numpy(lr_t), torch.from_numpy(lr_p), torch.from_numpy(hr))   # [n,51,h,w]
    got = ops.fused_frame_input(_cuda(lr_t), _cuda(lr_p), _cuda(hr))                             # [n,h,w,64] bf16
    assert got.shape == (n, h, w, 64) and got.dtype == torch.bfloat16
    got = got.float().cpu().permute(0, 3, 1, 2)
    assert got[:, 51:].abs().max().item() == 0.0
    # output is stored as bf16: compare against the bf16-rounded oracle; a rare fp16 grid-ulp flip
    #  moves a sample by one grid step -> allow a few outliers
    want_bf = want.to(torch.bfloat16).float()
    diff = (got[:, :51] - want_bf).abs()
    assert (diff > 1e-2).float().mean().item() < 2e-3
    assert torch.equal(got[:, :3], want_bf[:, :3])
    # first frame: zeros state (main.py:191-193)
    first = ops.fused_frame_input(_cuda(lr_t)).float().cpu().permute(0, 3, 1, 2)
    assert first[:, 3:].abs().max().item() == 0.0 and torch.equal(first[:, :3], want_bf[:, :3])


def _motion_grid(n, ho, wo, amp_px, seed):
    """identity sampling grid + a displacement of up to +-amp_px pixels (what an optical-flow field looks like)."""
    ys = (np.arange(ho, dtype=np.float32) + 0.5) * 2 / ho - 1
    xs = (np.arange(wo, dtype=np.float32) + 0.5) * 2 / wo - 1
    g = np.stack(np.broadcast_arrays(xs[None, None, :], ys[None, :, None]), axis=-1).astype(np.float32)
    g = np.broadcast_to(g, (n, ho, wo, 2)).copy()
    d = synth.det_uniform((n, ho, wo, 2), seed, -1.0, 1.0)
    g[..., 0] += d[..., 0] * np.float32(2.0 * amp_px / wo)
    g[..., 1] += d[..., 1] * np.float32(2.0 * amp_px / ho)
    return g


@pytest.mark.parametrize("n,c,h,w,kind", [(2, 3, 96, 128, "motion"), (1, 3, 100, 200, "motion_edges"), (1, 3, 64, 64, "random"),
                                          (2, 2, 72, 132, "mixed"), (1, 3, 40, 68, "all_out")])
def test_warp_on_motion_and_mixed_fields_vs_oracle(n, c, h, w, kind):
    """The warp kernel against the numpy oracle and torch.grid_sample, <= 1e-5 (BASELINE.json north_star), on the kinds of
    sampling field it meets: a motion-like field (identity +- 2 px), one that pushes taps across every image edge, a
    random field, a half-and-half field and one that leaves the image entirely (zeros padding)."""
    from tecogan_b200 import ops
    img = synth.det_uniform((n, c, h, w), 21, -1, 1)
    if kind == "motion":
        grid = _motion_grid(n, h, w, 2.0, 22)
    elif kind == "motion_edges":
        grid = _motion_grid(n, h, w, 3.0, 23) * np.float32(1.04)            # the outer ring samples outside the image
    elif kind == "random":
        grid = synth.det_uniform((n, h, w, 2), 24, -1.1, 1.1)
    elif kind == "mixed":
        grid = _motion_grid(n, h, w, 1.5, 25)
        grid[:, :, w // 2:] = synth.det_uniform((n, h, w - w // 2, 2), 26, -1.0, 1.0)
    else:
        grid = synth.det_uniform((n, h, w, 2), 27, 1.5, 3.0)
    want = glue_np.warp(img, grid)
    got = ops.warp(_cuda(img), _cuda(grid)).cpu().numpy()
    assert np.abs(got - want).max() <= 1e-5
    want_t = O.warp(torch.from_numpy(img), torch.from_numpy(grid)).numpy()
    assert np.abs(got - want_t).max() <= 1e-5
    if kind == "all_out":
        assert np.abs(got).max() == 0.0

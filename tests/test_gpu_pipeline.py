"""GPU parity of the end-to-end API that bench.py's `e2e` measures: tecogan_b200.pipeline.ClipPipeline (the loop of the
reference's inference driver, main.py:173-220, kept on the device) against the CPU oracle and against G.infer_clip, for
every output format (f32 = the parity default; fp16 = what the reference's autocast path emits, main.py:171-172; uint8
NHWC = what save_as_gif makes of it, code/ops.py:234-237)."""
import math
import types

import pytest
import torch

from oracle import synth, tecogan_oracle as O

pytestmark = pytest.mark.gpu


def _make(gain=1.0, nres=16):
    from tecogan_b200 import models
    ref = O.OracleGenerator(3, nres).eval()
    O.load_numpy_state(ref, synth.fill_state_dict(ref.state_dict(), seed=1, gain=gain))
    G = models.generator(3, types.SimpleNamespace(num_resblock=nres))
    G.load_state_dict(ref.state_dict())
    return ref, G.cuda().eval()


def _psnr(a, b):
    mse = ((a.double() - b.double()) ** 2).mean().item()
    return 99.0 if mse == 0 else 10.0 * math.log10(1.0 / mse)


def test_run_device_and_run_host_vs_oracle_cfg1():
    """BASELINE cfg1 (10-frame 64x64 LR clip -> 256x256) through both entry points of the pipeline vs O.infer_clip."""
    from tecogan_b200.pipeline import ClipPipeline
    torch.set_num_threads(8)
    ref, G = _make(1.0)
    b, t, h, w = 1, 10, 64, 64
    r = torch.from_numpy(synth.clip_inputs(b, t, h, w, seed=1234, hi=0.25))
    want = O.infer_clip(ref, r)                                           # [B,T,3,4h,4w]
    pipe = ClipPipeline(G, b, t, h, w)
    dev = pipe.run_device(r.cuda()).cpu()
    assert dev.shape == want.shape
    assert _psnr(dev, want) >= 50.0 and (dev - want).abs().max().item() <= 1e-2
    lr_host = r.clone().pin_memory()
    out_host = torch.empty((t, b, 3, 4 * h, 4 * w), dtype=torch.float32).pin_memory()
    pipe.run_host(lr_host, out_host)
    host = out_host.transpose(0, 1)                                       # frame-major host layout -> [B,T,...]
    assert _psnr(host, want) >= 50.0 and (host - want).abs().max().item() <= 1e-2
    per = [_psnr(host[:, k], want[:, k]) for k in range(t)]
    assert min(per) >= 50.0, per
    assert torch.equal(host, dev)                                         # the two entry points run the same kernels


def test_pipeline_is_bit_equal_to_infer_clip_at_720p():
    """2 clips x 3 frames of 320x180: run_device == G.infer_clip bit for bit; run_host writes frame f of clip c at
    out_host[f, c] (frame-major: one contiguous D2H transfer per frame) with the same bits; a second run over the same
    pipeline (workspace, staging buffers, copy stream reused) reproduces them."""
    from tecogan_b200.pipeline import ClipPipeline
    _, G = _make(1.7)
    b, t, h, w = 2, 3, 180, 320
    r = torch.from_numpy(synth.clip_inputs(b, t, h, w, seed=77, hi=0.25))
    want = G.infer_clip(r.cuda()).cpu()
    pipe = ClipPipeline(G, b, t, h, w)
    assert torch.equal(pipe.run_device(r.cuda()).cpu(), want)
    lr_host = r.clone().pin_memory()
    out_host = torch.full((t, b, 3, 4 * h, 4 * w), -1.0, dtype=torch.float32).pin_memory()
    for _ in range(2):
        pipe.run_host(lr_host, out_host)
        for f in range(t):
            for c in range(b):
                assert torch.equal(out_host[f, c], want[c, f]), (f, c)
    bi, bo = pipe.bytes_per_run()
    assert bi == r.numel() * 4 and bo == want.numel() * 4


@pytest.mark.parametrize("fmt", ["f16", "u8"])
def test_compact_output_formats(fmt):
    """out_dtype variants of run_host: fp16 planar frames == the f32 result rounded to fp16 (what the reference's autocast
    path emits, main.py:171-172,214); uint8 NHWC frames == (x * 255).astype(uint8) of the f32 result in [T,B,H,W,3] order,
    i.e. exactly what save_as_gif builds before writing (code/ops.py:234-237: transpose (0,2,3,1), * 255, astype(uint8)).
    Bit-exact against the f32 path, and the recurrence (which keeps consuming the f32 frames) is unchanged."""
    from tecogan_b200.pipeline import ClipPipeline
    _, G = _make(1.7)
    b, t, h, w = 2, 3, 36, 52
    r = torch.from_numpy(synth.clip_inputs(b, t, h, w, seed=78, hi=0.25))
    want = G.infer_clip(r.cuda()).cpu().transpose(0, 1).contiguous()      # [T,B,3,H,W] f32
    pipe = ClipPipeline(G, b, t, h, w)
    lr_host = r.clone().pin_memory()
    if fmt == "f16":
        out = torch.empty((t, b, 3, 4 * h, 4 * w), dtype=torch.float16).pin_memory()
        pipe.run_host(lr_host, out, out_dtype=torch.float16)
        assert torch.equal(out, want.half())
        assert pipe.bytes_per_run(torch.float16)[1] == want.numel() * 2
    else:
        out = torch.empty((t, b, 4 * h, 4 * w, 3), dtype=torch.uint8).pin_memory()
        pipe.run_host(lr_host, out, out_dtype=torch.uint8)
        ref8 = (want.permute(0, 1, 3, 4, 2) * 255).to(torch.uint8)        # numpy astype(uint8) truncates like torch's cast
        assert torch.equal(out, ref8)
        assert pipe.bytes_per_run(torch.uint8)[1] == want.numel()


def test_pipeline_rejects_wrong_shapes_and_devices():
    from tecogan_b200.pipeline import ClipPipeline
    _, G = _make(1.0, nres=1)
    pipe = ClipPipeline(G, 1, 2, 16, 16)
    with pytest.raises(RuntimeError):
        pipe.run_device(torch.zeros(1, 2, 3, 16, 16))                     # CPU tensor: no fallback
    with pytest.raises((RuntimeError, AssertionError)):
        pipe.run_device(torch.zeros(1, 3, 3, 16, 16, device="cuda"))      # wrong frame count

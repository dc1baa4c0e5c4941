"""Quick per-layer timing probe (CUDA events) for the conv kernels at cfg2 (720p) sizes.
Not the bench; prints achieved TFLOP/s per layer type so kernel work can be prioritised."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "pytorch-tecogan_b200"))
from tecogan_b200 import _native as nt  # noqa: E402

lib = nt.lib()
AMODE = int(os.environ.get("TG_AMODE", "0"))


def pack(kind, cin, cout):
    w = (torch.rand(cout if kind == 0 else cin, cin if kind == 0 else cout, 3, 3, device="cuda") - 0.5) * 0.1
    b = torch.rand(cout, device="cuda")
    p = torch.zeros(lib.tg_packed_conv_bytes(kind, cin, cout), dtype=torch.uint8, device="cuda")
    nt.check(lib.tg_pack_weights(kind, nt.ptr(w), nt.ptr(b), cin, cout, nt.ptr(p), nt.stream_ptr()))
    return p


def timeit(fn, iters=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record()
    for _ in range(iters):
        fn()
    e.record()
    torch.cuda.synchronize()
    return s.elapsed_time(e) / iters * 1e-3


def layer(kind, n, h, w, cin, cout, name):
    x = (torch.rand(n, h, w, cin, device="cuda") - 0.5).to(torch.bfloat16)
    p = pack(kind, cin, cout)
    s = 2 if kind == 1 else 1
    y = torch.empty(n, h * s, w * s, cout, dtype=torch.bfloat16, device="cuda")
    if kind == 0:
        fn = lambda: nt.check(lib.tg_conv3x3_fwd(nt.ptr(x), nt.ptr(p), None, nt.ptr(y), n, h, w, cin, cout, 1, AMODE, nt.stream_ptr()))
    else:
        fn = lambda: nt.check(lib.tg_convT3x3s2_fwd(nt.ptr(x), nt.ptr(p), nt.ptr(y), n, h, w, cin, cout, 1, AMODE, nt.stream_ptr()))
    t = timeit(fn)
    fl = 2.0 * 9 * cin * cout * n * h * w
    print(f"{name:34s} {t*1e6:9.1f} us  {fl/t/1e12:8.1f} TFLOP/s  in+out {(x.numel()+y.numel())*2/t/1e9:7.0f} GB/s", flush=True)


if __name__ == "__main__":
    N = int(os.environ.get("TG_N", "1"))
    H, W = 180, 320
    print("amode", AMODE, "N", N)
    layer(0, N, H, W, 64, 64, "conv 64->64 @1x (trunk)")
    layer(1, N, H, W, 64, 64, "convT 64->64 @1x->2x")
    layer(0, N, 2 * H, 2 * W, 64, 64, "conv 64->64 @2x")
    layer(0, N, 2 * H, 2 * W, 64, 128, "conv 64->128 @2x")
    layer(0, N, 2 * H, 2 * W, 128, 128, "conv 128->128 @2x")
    layer(1, N, 2 * H, 2 * W, 128, 128, "convT 128->128 @2x->4x")
    layer(0, N, 4 * H, 4 * W, 128, 64, "conv 128->64 @4x")
    # output conv
    x = (torch.rand(N, 4 * H, 4 * W, 64, device="cuda") - 0.5).to(torch.bfloat16)
    p = pack(0, 64, 3)
    out = torch.empty(N, 3, 4 * H, 4 * W, device="cuda")
    t = timeit(lambda: nt.check(lib.tg_conv3x3_out_sigmoid(nt.ptr(x), nt.ptr(p), nt.ptr(out), None, N, 4 * H, 4 * W, AMODE, nt.stream_ptr())))
    print(f"{'conv 64->3 + sigmoid @4x':34s} {t*1e6:9.1f} us  {(x.numel()*2+out.numel()*4)/t/1e9:7.0f} GB/s")
    # whole generator frame
    nres = 16
    flat = (torch.rand(lib.tg_gen_param_count(nres), device="cuda") - 0.5) * 0.05
    packed = torch.zeros(lib.tg_gen_packed_bytes(nres), dtype=torch.uint8, device="cuda")
    nt.check(lib.tg_gen_pack(nt.ptr(flat), nres, nt.ptr(packed), nt.stream_ptr()))
    ws = torch.empty(lib.tg_gen_workspace_bytes(N, H, W), dtype=torch.uint8, device="cuda")
    x0 = torch.rand(N, H, W, 64, device="cuda").to(torch.bfloat16)
    for am, nm in ((AMODE, "per-layer launches"), (2, "frame kernel")):
        t = timeit(lambda: nt.check(lib.tg_gen_forward(nt.ptr(packed), nres, nt.ptr(x0), nt.ptr(out), None, nt.ptr(ws), ws.numel(), N, H, W, am, nt.stream_ptr())), iters=10)
        print(f"generator frame 720p ({nm}): {t*1e3:.3f} ms  -> {N/t:.1f} frames/s  {N*486.45e9/t/1e12:.1f} TFLOP/s", flush=True)
    lr = torch.rand(N, 3, H, W, device="cuda")
    hr = torch.rand(N, 3, 4 * H, 4 * W, device="cuda")
    t = timeit(lambda: nt.check(lib.tg_fused_warp_s2d_concat(nt.ptr(lr), nt.ptr(lr), nt.ptr(hr), nt.ptr(x0), N, H, W, 3 * H * W, 48 * H * W, nt.stream_ptr())))
    print(f"fused warp+s2d+concat 720p: {t*1e6:.1f} us  {N*16*H*W*18.4/t/1e9:.0f} GB/s (18.4 B/HR px)")

"""Device-resident clip pipeline — the loop of the reference's inference driver
(main.py:173-219) without its three PCIe hops per frame (main.py:195,203,214) and without the
CPU-side upscale_four (main.py:186).

    pipe = ClipPipeline(G, batch=2, frames=100, h=180, w=320)
    hr = pipe.run_device(lr_cuda)                 # [B,T,3,h,w] cuda f32 -> [B,T,3,4h,4w] cuda f32
    pipe.run_host(lr_pinned, out_pinned)          # host buffers in, host buffers out (H2D/D2H overlapped)
    pipe.run_host(lr_pinned, out_u8, out_dtype=torch.uint8)   # compact frames for the encoder side

The frame recurrence is strictly sequential; the B clips of a batch are independent and are
processed together so that every kernel launch has B x more tiles to spread over the 148 SMs.

Output formats of ``run_host`` (what crosses PCIe; the recurrence itself always runs on the f32 estimate kept in
the workspace):
  torch.float32  [T,B,3,4h,4w]  the parity default, bit-identical to ``generator.infer_clip``
  torch.float16  [T,B,3,4h,4w]  the f32 frames rounded to fp16 — what the reference's GPU path emits under autocast
                                (main.py:171-172,214); half the D2H bytes
  torch.uint8    [T,B,4h,4w,3]  ``(x * 255).astype(uint8)`` in the NHWC order ``save_as_gif`` writes
                                (code/ops.py:234-237); a quarter of the D2H bytes
"""
import ctypes

import torch

from . import _native as _nt

_FMT = {torch.float32: (_nt.OUT_F32, 4), torch.float16: (_nt.OUT_F16, 2), torch.uint8: (_nt.OUT_U8, 1)}


class ClipPipeline:
    def __init__(self, gen, batch, frames, h, w, device=None):
        self.gen = gen
        self.b, self.t, self.h, self.w = int(batch), int(frames), int(h), int(w)
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        if self.dev.type != "cuda":
            raise RuntimeError("ClipPipeline needs a CUDA device (no CPU fallback)")
        with torch.cuda.device(self.dev):
            lib = _nt.lib()
            self.nres = int(gen.num)
            self.ws = torch.empty(lib.tg_gen_workspace_bytes(self.b, self.h, self.w), dtype=torch.uint8, device=self.dev)
        self._stage = {}                # out dtype -> [T,B,...] device staging for run_host
        self._lr_dev = None
        self._copy_stream = None

    def _out_shape(self, dtype):
        t, b, h, w = self.t, self.b, self.h, self.w
        return (t, b, 4 * h, 4 * w, 3) if dtype == torch.uint8 else (t, b, 3, 4 * h, 4 * w)

    # ------------------------------------------------------------------ device-resident inputs
    @torch.no_grad()
    def run_device(self, lr, out=None):
        """lr [B,T,3,h,w] f32 on the device -> out [B,T,3,4h,4w] f32 (one C-ABI call per clip batch)."""
        lr = _nt.require_cuda_f32(lr, "run_device(lr)")
        if tuple(lr.shape) != (self.b, self.t, 3, self.h, self.w) or lr.device != self.dev:
            raise RuntimeError(f"run_device: expected {(self.b, self.t, 3, self.h, self.w)} on {self.dev}, got "
                               f"{tuple(lr.shape)} on {lr.device}")
        with torch.cuda.device(self.dev):
            lib = _nt.lib()
            if out is None:
                out = torch.empty((self.b, self.t, 3, 4 * self.h, 4 * self.w), dtype=torch.float32, device=self.dev)
            packed = self.gen.packed_weights()
            _nt.check(lib.tg_gen_clip_forward(_nt.ptr(packed), self.nres, _nt.ptr(lr), _nt.ptr(out), _nt.ptr(self.ws),
                                              self.ws.numel(), self.b, self.t, self.h, self.w, int(self.gen.amode),
                                              _nt.stream_ptr(self.dev)))
        return out

    # --------------------------------------------------------------------- host-resident inputs
    @torch.no_grad()
    def run_host(self, lr_host, out_host, out_dtype=torch.float32):
        """End-to-end call with HOST buffers: lr_host [B,T,3,h,w] f32 (pinned) is copied to the device,
        every finished HR frame is copied back into out_host (pinned, FRAME-major so that each frame is one
        contiguous transfer; shape / dtype per ``out_dtype``, see the module docstring) on a side stream while
        the next frame computes.  Returns when all frames are in host memory."""
        if out_dtype not in _FMT:
            raise RuntimeError(f"run_host: out_dtype must be float32, float16 or uint8 (got {out_dtype})")
        fmt, esize = _FMT[out_dtype]
        b, t, h, w = self.b, self.t, self.h, self.w
        if tuple(lr_host.shape) != (b, t, 3, h, w) or lr_host.dtype != torch.float32:
            raise RuntimeError(f"run_host: lr_host must be f32 {(b, t, 3, h, w)}, got {lr_host.dtype} {tuple(lr_host.shape)}")
        if tuple(out_host.shape) != self._out_shape(out_dtype) or out_host.dtype != out_dtype or not out_host.is_contiguous():
            raise RuntimeError(f"run_host: out_host must be contiguous {out_dtype} {self._out_shape(out_dtype)}, got "
                               f"{out_host.dtype} {tuple(out_host.shape)}")
        if fmt != _nt.OUT_F32 and int(self.gen.amode) != _nt.AMODE_FRAME:
            raise RuntimeError("run_host: fp16 / uint8 outputs are written by the frame kernel's output conv (amode FRAME)")
        with torch.cuda.device(self.dev):
            lib = _nt.lib()
            if out_dtype not in self._stage:
                self._stage[out_dtype] = torch.empty(self._out_shape(out_dtype), dtype=out_dtype, device=self.dev)
            if self._lr_dev is None:
                self._lr_dev = torch.empty((b, t, 3, h, w), dtype=torch.float32, device=self.dev)
                self._copy_stream = torch.cuda.Stream(device=self.dev)
            frames = self._stage[out_dtype]
            main = torch.cuda.current_stream(self.dev)
            self._lr_dev.copy_(lr_host, non_blocking=True)                       # H2D, on the compute stream
            packed = self.gen.packed_weights()
            lr_frame, hr_frame = 3 * h * w, 48 * h * w                           # elements per clip and frame
            lr_ptr, fr_ptr = self._lr_dev.data_ptr(), frames.data_ptr()
            vp = ctypes.c_void_p
            stream = _nt.stream_ptr(self.dev)
            for f in range(t):
                lr_t = vp(lr_ptr + 4 * f * lr_frame)
                lr_prev = vp(lr_ptr + 4 * (f - 1) * lr_frame) if f else vp(0)
                cur_hr = vp(fr_ptr + esize * f * b * hr_frame)
                if fmt == _nt.OUT_F32 and int(self.gen.amode) != _nt.AMODE_FRAME:
                    # per-layer mode (measurement): no interleaved copy, plain step on the planar f32 frames
                    prev_hr = vp(fr_ptr + 4 * (f - 1) * b * hr_frame) if f else vp(0)
                    _nt.check(lib.tg_gen_clip_step(_nt.ptr(packed), self.nres, lr_t, lr_prev, prev_hr, cur_hr, _nt.ptr(self.ws),
                                                   self.ws.numel(), b, h, w, t * lr_frame, hr_frame, hr_frame,
                                                   int(self.gen.amode), stream))
                else:
                    # chained step: frame f-1's estimate is gathered from the workspace's interleaved f32 copy
                    # (bit-identical to the plain step, 3x fewer L1 sectors per tap); only cur_hr leaves the GPU
                    _nt.check(lib.tg_gen_clip_step_fmt(_nt.ptr(packed), self.nres, lr_t, lr_prev, 1 if f == 0 else 0, cur_hr, fmt,
                                                       _nt.ptr(self.ws), self.ws.numel(), b, h, w, t * lr_frame, hr_frame, stream))
                ev = torch.cuda.Event()
                ev.record(main)
                self._copy_stream.wait_event(ev)
                with torch.cuda.stream(self._copy_stream):
                    out_host[f].copy_(frames[f], non_blocking=True)              # D2H overlaps frame f+1
            self._copy_stream.synchronize()
            main.synchronize()
        return out_host

    def bytes_per_run(self, out_dtype=torch.float32):
        """(host->device, device->host) bytes of one run_host call."""
        b, t, h, w = self.b, self.t, self.h, self.w
        return b * t * 3 * h * w * 4, b * t * 48 * h * w * _FMT[out_dtype][1]

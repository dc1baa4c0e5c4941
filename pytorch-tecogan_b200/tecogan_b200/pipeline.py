"""Device-resident clip pipeline — the loop of the reference's inference driver
(main.py:173-219) without its three PCIe hops per frame (main.py:195,203,214) and without the
CPU-side upscale_four (main.py:186).

    pipe = ClipPipeline(G, batch=2, frames=100, h=180, w=320)
    hr = pipe.run_device(lr_cuda)                 # [B,T,3,h,w] cuda f32 -> [B,T,3,4h,4w] cuda f32
    pipe.run_host(lr_pinned, out_pinned)          # host buffers in, host buffers out (H2D/D2H overlapped)

The frame recurrence is strictly sequential; the B clips of a batch are independent and are
processed together so that every kernel launch has B x more tiles to spread over the 148 SMs.
"""
import torch

from . import _native as _nt


class ClipPipeline:
    def __init__(self, gen, batch, frames, h, w, device=None):
        self.gen = gen
        self.b, self.t, self.h, self.w = int(batch), int(frames), int(h), int(w)
        self.dev = torch.device(device if device is not None else f"cuda:{torch.cuda.current_device()}")
        lib = _nt.lib()
        self.nres = int(gen.num)
        self.ws = torch.empty(lib.tg_gen_workspace_bytes(self.b, self.h, self.w), dtype=torch.uint8, device=self.dev)
        self._frames_tb = None          # [T,B,3,4h,4w] device staging for run_host
        self._lr_dev = None
        self._copy_stream = None

    # ------------------------------------------------------------------ device-resident inputs
    @torch.no_grad()
    def run_device(self, lr, out=None):
        """lr [B,T,3,h,w] f32 on the device -> out [B,T,3,4h,4w] f32 (one C-ABI call per clip batch)."""
        lib = _nt.lib()
        lr = _nt.require_cuda_f32(lr, "run_device(lr)")
        assert tuple(lr.shape) == (self.b, self.t, 3, self.h, self.w), tuple(lr.shape)
        if out is None:
            out = torch.empty((self.b, self.t, 3, 4 * self.h, 4 * self.w), dtype=torch.float32, device=self.dev)
        packed = self.gen.packed_weights()
        _nt.check(lib.tg_gen_clip_forward(_nt.ptr(packed), self.nres, _nt.ptr(lr), _nt.ptr(out), _nt.ptr(self.ws),
                                          self.ws.numel(), self.b, self.t, self.h, self.w, int(self.gen.amode),
                                          _nt.stream_ptr()))
        return out

    # --------------------------------------------------------------------- host-resident inputs
    @torch.no_grad()
    def run_host(self, lr_host, out_host):
        """End-to-end call with HOST buffers: lr_host [B,T,3,h,w] f32 (pinned) is copied to the device,
        every finished HR frame is copied back into out_host [T,B,3,4h,4w] f32 (pinned, frame-major so
        that each frame is one contiguous transfer) on a side stream while the next frame computes.
        Returns when all frames are in host memory."""
        lib = _nt.lib()
        b, t, h, w = self.b, self.t, self.h, self.w
        assert tuple(lr_host.shape) == (b, t, 3, h, w) and tuple(out_host.shape) == (t, b, 3, 4 * h, 4 * w)
        if self._frames_tb is None:
            self._frames_tb = torch.empty((t, b, 3, 4 * h, 4 * w), dtype=torch.float32, device=self.dev)
            self._lr_dev = torch.empty((b, t, 3, h, w), dtype=torch.float32, device=self.dev)
            self._copy_stream = torch.cuda.Stream(device=self.dev)
        main = torch.cuda.current_stream(self.dev)
        self._lr_dev.copy_(lr_host, non_blocking=True)                       # H2D, on the compute stream
        packed = self.gen.packed_weights()
        lr_frame, hr_frame = 3 * h * w, 48 * h * w
        lr_ptr, fr_ptr = self._lr_dev.data_ptr(), self._frames_tb.data_ptr()
        import ctypes
        vp = ctypes.c_void_p
        for f in range(t):
            lr_t = vp(lr_ptr + 4 * f * lr_frame)
            lr_prev = vp(lr_ptr + 4 * (f - 1) * lr_frame) if f else vp(0)
            prev_hr = vp(fr_ptr + 4 * (f - 1) * b * hr_frame) if f else vp(0)
            cur_hr = vp(fr_ptr + 4 * f * b * hr_frame)
            # frame f-1's output is untouched between the two calls: the chained step may gather from the workspace's
            # interleaved copy of it (bit-identical, 3x fewer L1 sectors per tap)
            step = lib.tg_gen_clip_step_chained if f else lib.tg_gen_clip_step
            _nt.check(step(_nt.ptr(packed), self.nres, lr_t, lr_prev, prev_hr, cur_hr, _nt.ptr(self.ws),
                           self.ws.numel(), b, h, w, t * lr_frame, hr_frame, hr_frame,
                           int(self.gen.amode), _nt.stream_ptr()))
            ev = torch.cuda.Event()
            ev.record(main)
            self._copy_stream.wait_event(ev)
            with torch.cuda.stream(self._copy_stream):
                out_host[f].copy_(self._frames_tb[f], non_blocking=True)     # D2H overlaps frame f+1
        self._copy_stream.synchronize()
        main.synchronize()
        return out_host

    def bytes_per_run(self):
        b, t, h, w = self.b, self.t, self.h, self.w
        return b * t * 3 * h * w * 4, b * t * 48 * h * w * 4

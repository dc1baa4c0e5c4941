"""Data-parallel plumbing for the training step: one process per GPU, both networks replicated, the batch
of clips split across ranks, and ONE exchange per network and step — an all-reduce (mean) of its flat f32 gradient
bucket over NCCL / NVLink, issued right after that network's backward pass so that it overlaps the other network's
backward.  No other collective exists on the path (BatchNorm statistics stay per rank, as torch DDP's default).

The backward kernels ADD parameter gradients straight into a flat buffer laid out like the flat parameter vector
(include/tecogan_b200.h: tg_gen_backward / tg_disc_backward ``flat_grad``).  ``bind_flat_grads`` makes every
``p.grad`` a view of that buffer, so the optimizer (stock Adam, main.py:239-243) and GradScaler see ordinary gradients
while the collective moves one contiguous 7 MB (generator) / 13 MB (discriminator) message.
"""
import torch
import torch.distributed as dist


def bind_flat_grads(module):
    """Allocate (once) the module's flat gradient bucket and point every parameter's ``.grad`` into it."""
    params = [p for _, p in module.named_parameters()]
    bucket = getattr(module, "_grad_bucket", None)
    total = sum(p.numel() for p in params)
    if bucket is None or bucket.numel() != total or bucket.device != params[0].device:
        bucket = torch.zeros(total, dtype=torch.float32, device=params[0].device)
        module._grad_bucket = bucket
    o = 0
    for p in params:
        p.grad = bucket[o:o + p.numel()].view_as(p)
        o += p.numel()
    return bucket


def zero_flat_grads(module):
    """optimizer.zero_grad() for a bound module: clear the bucket and re-point the ``.grad`` views (the optimizer's own
    zero_grad(set_to_none=True) would drop them)."""
    bucket = bind_flat_grads(module)
    bucket.zero_()
    return bucket


def unbind_flat_grads(module):
    if getattr(module, "_grad_bucket", None) is not None:
        module._grad_bucket = None
        for p in module.parameters():
            p.grad = None


# measurement only (bench.py: exposed all-reduce time = step time with - step time without): skip the collective, the
# replicas then diverge
SKIP_ALLREDUCE = False


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


class GraphSegments:
    """A training step captured as a SEQUENCE of CUDA graphs with the NCCL calls between them left eager
    (tecogan_b200.train): ``eager(fn)`` is called at every collective / wait.  While a capture is running it closes the
    current graph, runs ``fn`` on the capturing stream, remembers it, and opens the next graph in the same memory pool;
    ``replay()`` then alternates graph replays and the remembered calls.  Outside a capture ``eager`` just calls ``fn``.
    (Capturing the collectives INSIDE one graph worked but hung torch.distributed's teardown - DESIGN.md section 6.)"""

    def __init__(self):
        self.ops = []                   # torch.cuda.CUDAGraph | callable, in execution order
        self.capturing = False
        self._graph = None
        self._pool = None
        self._mode = "thread_local"     # NCCL's watchdog thread polls events concurrently

    def begin(self):
        self.capturing = True
        self._open()

    def _open(self):
        g = torch.cuda.CUDAGraph()
        if self._pool is None:
            self._pool = torch.cuda.graph_pool_handle()      # one private memory pool for all segments of the step
        g.capture_begin(pool=self._pool, capture_error_mode=self._mode)
        self._graph = g

    def _close(self):
        self._graph.capture_end()
        self.ops.append(self._graph)
        self._graph = None

    def eager(self, fn):
        if not self.capturing:
            return fn()
        self._close()
        r = fn()
        self.ops.append(fn)
        self._open()
        return r

    def end(self):
        self._close()
        self.capturing = False

    def replay(self):
        for op in self.ops:
            if isinstance(op, torch.cuda.CUDAGraph):
                op.replay()
            else:
                op()


_segments = None                        # the GraphSegments of the step being captured / None (set by tecogan_b200.train)


def _eager(fn):
    return _segments.eager(fn) if _segments is not None else fn()


class GradSync:
    """Asynchronous mean all-reduce of one flat bucket.  ``start`` enqueues the collective behind the kernels already on
    the current stream (NCCL runs it on its own stream: later kernels of the current stream overlap it); ``finish``
    makes the current stream wait for it and applies the 1/world scaling.  Both NCCL-facing calls go through the step's
    GraphSegments (when one is capturing) so that they stay outside the CUDA graphs."""

    def __init__(self, group=None):
        self.group = group
        self.work = None
        self.bucket = None              # kept after finish(): a captured step replays _start / _wait on the same bucket
        self.active = False

    def _start(self):
        # (SKIP_ALLREDUCE is read at call time: a captured step replays this method)
        self.work = None if SKIP_ALLREDUCE else dist.all_reduce(self.bucket, op=dist.ReduceOp.SUM, group=self.group, async_op=True)

    def _wait(self):
        if self.work is not None:
            self.work.wait()
            self.work = None

    def start(self, bucket):
        if world_size() == 1 or (SKIP_ALLREDUCE and _segments is None):
            return
        self.bucket = bucket
        self.active = True
        _eager(self._start)

    def finish(self):
        if not self.active:
            return
        _eager(self._wait)
        self.bucket.mul_(1.0 / dist.get_world_size(self.group))
        self.active = False


def shard_batch(t, rank=None, world=None):
    """rank's contiguous share of the clip (batch) dimension; the global batch must divide evenly (cfg5: 32 clips)."""
    world = world_size() if world is None else world
    rank = (dist.get_rank() if world > 1 else 0) if rank is None else rank
    if t.shape[0] % world:
        raise RuntimeError(f"global batch {t.shape[0]} does not divide over {world} ranks")
    per = t.shape[0] // world
    return t[rank * per:(rank + 1) * per]


def broadcast_parameters(module, src=0):
    """make the replicas identical before the first step (parameters and BatchNorm buffers).  The in-place write goes
    through ``detach()`` (shares the parameter's version counter, unlike ``.data``) and the module's packed bf16 weight
    caches are dropped explicitly, so a broadcast after the first forward cannot leave stale packed weights behind."""
    if world_size() > 1:
        with torch.no_grad():
            for t in list(module.parameters()) + list(module.buffers()):
                dist.broadcast(t.detach(), src=src)
    if hasattr(module, "invalidate_packed"):
        module.invalidate_packed()


def bucket_bound(module, params=None):
    """True when the module's flat gradient bucket is live: every parameter's ``.grad`` IS the bucket view laid out by
    bind_flat_grads.  An ordinary backward after ``zero_grad(set_to_none=True)`` (p.grad None) is not: the backward
    kernels must then return gradients to autograd instead of adding into a bucket nobody reads."""
    bucket = getattr(module, "_grad_bucket", None)
    if bucket is None:
        return False
    params = [p for _, p in module.named_parameters()] if params is None else params
    o = bucket.data_ptr()
    for p in params:
        g = p.grad
        if g is None or g.data_ptr() != o or g.dtype != torch.float32:
            return False
        o += 4 * p.numel()
    return True

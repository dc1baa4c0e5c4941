"""CPU (gloo, world_size 2) tests of the data-parallel host logic of the training step (tecogan_b200.parallel):
flat gradient buckets, the asynchronous mean all-reduce, batch sharding and replica broadcast.  The collective on the
GPU box is the same torch.distributed call over NCCL."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from tecogan_b200 import parallel as P
        torch.manual_seed(100 + rank)                      # replicas start different ...
        m = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3), torch.nn.BatchNorm2d(4), torch.nn.Linear(5, 2))
        P.broadcast_parameters(m, src=0)                   # ... and are made identical
        flat0 = torch.cat([p.detach().flatten() for p in m.parameters()])
        gathered = [torch.zeros_like(flat0) for _ in range(world)]
        dist.all_gather(gathered, flat0)
        same = all(torch.equal(gathered[0], g) for g in gathered)
        bucket = P.bind_flat_grads(m)
        views_ok = all(p.grad.data_ptr() >= bucket.data_ptr() for p in m.parameters())
        # a "backward" that adds rank-dependent gradients in place, like the wgrad kernels do
        bucket += float(rank + 1)
        sync = P.GradSync()
        sync.start(bucket)
        sync.finish()
        mean_ok = bool(torch.allclose(bucket, torch.full_like(bucket, (1 + world) / 2)))
        grad_ok = all(torch.allclose(p.grad, torch.full_like(p, (1 + world) / 2)) for p in m.parameters())
        opt = torch.optim.Adam(m.parameters(), 1e-3)
        opt.step()                                         # the optimizer consumes the views
        opt.zero_grad()                                    # drops them (set_to_none) ...
        b2 = P.zero_flat_grads(m)                          # ... re-bound and cleared for the next step
        rebound = b2.data_ptr() == bucket.data_ptr() and float(b2.abs().sum()) == 0.0 and \
            all(p.grad is not None for p in m.parameters())
        x = torch.arange(8 * 3).reshape(8, 3)
        shard = P.shard_batch(x)
        shard_ok = shard.shape[0] == 8 // world and int(shard[0, 0]) == rank * (8 // world) * 3
        try:
            P.shard_batch(torch.zeros(7, 1))
            uneven = False
        except RuntimeError:
            uneven = True
        q.put((rank, same, views_ok, mean_ok, grad_ok, rebound, shard_ok, uneven))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    for r in res:
        assert all(r[1:]), r


def test_single_process_is_a_no_op():
    from tecogan_b200 import parallel as P
    m = torch.nn.Linear(3, 2)
    bucket = P.bind_flat_grads(m)
    bucket += 2.0
    s = P.GradSync()
    s.start(bucket)
    s.finish()
    assert torch.equal(m.weight.grad, torch.full_like(m.weight, 2.0))
    assert P.world_size() == 1 and P.shard_batch(torch.zeros(4, 2)).shape[0] == 4
    P.unbind_flat_grads(m)
    assert m.weight.grad is None


def test_graph_segments_sequence_the_eager_calls_between_graphs(monkeypatch):
    """GraphSegments (the data-parallel step as graph segments with the NCCL calls eager between them): outside a capture
    eager() just calls; during a capture every eager() closes the running segment, runs and records the call, opens the
    next one in the same pool; replay() alternates them in order.  (The CUDA capture itself: tests/test_gpu_dp.py.)"""
    from tecogan_b200 import parallel as P
    log = []

    class FakeGraph:
        n = 0

        def __init__(self):
            FakeGraph.n += 1
            self.id = FakeGraph.n

        def capture_begin(self, pool=None, capture_error_mode=None):
            log.append(("begin", self.id, pool, capture_error_mode))

        def capture_end(self):
            log.append(("end", self.id))

        def replay(self):
            log.append(("replay", self.id))

    monkeypatch.setattr(torch.cuda, "CUDAGraph", FakeGraph)
    monkeypatch.setattr(torch.cuda, "graph_pool_handle", lambda: "pool0")
    seg = P.GraphSegments()
    assert seg.eager(lambda: 7) == 7 and log == []          # not capturing: a plain call
    seg.begin()
    assert seg.eager(lambda: log.append("allreduce G") or "w") == "w"
    seg.eager(lambda: log.append("wait G"))
    seg.end()
    assert log == [("begin", 1, "pool0", "thread_local"), ("end", 1), "allreduce G", ("begin", 2, "pool0", "thread_local"),
                   ("end", 2), "wait G", ("begin", 3, "pool0", "thread_local"), ("end", 3)]
    del log[:]
    seg.replay()
    assert log == [("replay", 1), "allreduce G", ("replay", 2), "wait G", ("replay", 3)]
    # GradSync routes its NCCL-facing calls through the step's segments only when one is installed
    assert P._segments is None and P._eager(lambda: 3) == 3


def test_train_mirror_interface():
    """the reference-facing surface of tecogan_b200.train (code/train.py): names, signature, namedtuple fields."""
    import inspect
    from tecogan_b200 import train as T
    assert T.Network._fields == ('gen_output', 'learning_rate', 'update_list', 'update_list_name', 'update_list_avg',
                                 'global_step', 'd_loss', 'gen_loss', 'fnet_loss', 'tb', 'target')
    assert list(inspect.signature(T.FRVSR_Train).parameters) == [
        'r_inputs', 'r_targets', 'args', 'discriminator_F', 'generator_F', 'step', 'counter1', 'counter2', 'optimizer_g',
        'optimizer_d']
    assert list(inspect.signature(T.TecoGAN).parameters)[:5] == ['r_inputs', 'r_targets', 'discriminator_F', 'generator_F', 'args']
    e = T.EMA(0.99)
    e.register("a", torch.zeros(()))
    assert abs(float(e("a", torch.tensor(2.0))) - 1.98) < 1e-6
    with pytest.raises(NotImplementedError):
        T.VGG19_slim(None, None)
    with pytest.raises(RuntimeError):                      # no CUDA device / CPU tensors: fails loudly, no fallback
        T.TecoGAN(torch.zeros(1, 10, 3, 32, 32), torch.zeros(1, 10, 3, 128, 128), None, None, None, 0, 0, 0, None, None)

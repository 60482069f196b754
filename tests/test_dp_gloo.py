"""world_size-2 gloo tests (CPU) of the data-parallel gradient exchange (csm/training/dp.py)."""
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


def _worker(rank, world, port, bucket_bytes, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "csm-train-pytorch_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from csm.training import dp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
        model[0].bias.requires_grad_(False)                   # a frozen parameter must be ignored
        sync = dp.GradSynchronizer(model.parameters(), bucket_bytes=bucket_bytes)
        for step in range(2):                                 # two steps: hooks/buckets must reset
            x = torch.full((3, 8), float(rank + 1 + step))
            model(x).sum().backward()
            sync.finish()
            local = [p.grad.clone() for p in model.parameters() if p.requires_grad]
            # reference: average of both ranks' gradients computed locally
            ref_model = torch.nn.Sequential(torch.nn.Linear(8, 16), torch.nn.ReLU(), torch.nn.Linear(16, 4))
            ref_model.load_state_dict(model.state_dict())
            acc = None
            for r in range(world):
                ref_model.zero_grad()
                ref_model(torch.full((3, 8), float(r + 1 + step))).sum().backward()
                gs = [p.grad.clone() for n, p in ref_model.named_parameters() if n != "0.bias"]
                acc = gs if acc is None else [a + g for a, g in zip(acc, gs)]
            for got, want in zip(local, acc):
                assert torch.allclose(got, want / world, atol=1e-6)
            model.zero_grad()
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("bucket_bytes", [None, 256])
def test_grad_synchronizer_world2(bucket_bytes):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, bucket_bytes, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _worker_deliver(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "csm-train-pytorch_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from csm.training import dp
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a = torch.nn.Parameter(torch.zeros(4, 3))
        b = torch.nn.Parameter(torch.zeros(5))
        c = torch.nn.Parameter(torch.zeros(2, 2))
        sync = dp.GradSynchronizer([a, b, c], bucket_bytes=32)
        assert sync.bucketed and len(sync._buckets) >= 2
        for step in range(2):
            # accumulation window of two micro-batches: first accumulates locally, second exchanges
            sync.accumulating = True
            assert sync.deliver(a, torch.full((4, 3), 1.0 + rank))          # hand-written backward path
            (b * (2.0 + rank)).sum().backward()                              # autograd path (hook)
            sync.finish()                                                    # no-op while accumulating
            sync.accumulating = False
            assert sync.deliver(a, torch.full((4, 3), 10.0 * (rank + 1)))
            (b * (3.0 + rank)).sum().backward()
            sync.finish()                                                    # c never got a gradient: counts as zero
            mean_a = sum((1.0 + r) + 10.0 * (r + 1) for r in range(world)) / world
            mean_b = sum((2.0 + r) + (3.0 + r) for r in range(world)) / world
            assert torch.allclose(a.grad, torch.full((4, 3), mean_a)), a.grad
            assert torch.allclose(b.grad, torch.full((5,), mean_b)), b.grad
            assert torch.equal(c.grad, torch.zeros(2, 2))
            for p in (a, b, c):
                p.grad = None
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_bucket_views_deliver_and_accumulation_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_deliver, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_single_process_is_noop():
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "csm-train-pytorch_b200"))
    from csm.training import dp
    lin = torch.nn.Linear(4, 4)
    lin(torch.ones(2, 4)).sum().backward()
    g = lin.weight.grad.clone()
    dp.GradSynchronizer(lin.parameters()).finish()
    assert torch.equal(lin.weight.grad, g)


def _worker_ragged(rank, world, port, q):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "csm-train-pytorch_b200"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    from csm.training import dp
    from csm.training.trainer import iterate_batches
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        # (1) ragged shapes: rank 0 holds [2, 5] frames, rank 1 [2, 7]; both pad their frame lists to the SAME fixed
        #     capacity (batch x capacity sequence length) without talking to each other; the padding is masked out
        #     with a zero upstream gradient, so the scatter kernel ignores it
        sync = dp.GradSynchronizer([torch.nn.Parameter(torch.zeros(3))], bucket_bytes=64)
        sync.text_capacity_seq = 8
        B, S = (2, 5) if rank == 0 else (2, 7)
        tok = torch.arange(B * S * 3).view(B, S, 3) + 1
        msk = torch.ones(B, S, 3, dtype=torch.uint8)
        dh = torch.ones(B, S, 4)
        t2, m2, d2 = sync.pad_to_capacity(tok, msk, dh)
        assert t2.shape == (16, 3) and m2.shape == (16, 3) and d2.shape == (16, 4)
        assert torch.equal(t2[:B * S], tok.view(-1, 3)) and torch.equal(m2[:B * S], msk.view(-1, 3))
        assert int(m2.sum()) == B * S * 3 and float(d2.sum()) == B * S * 4      # everything added is zero / masked
        gathered = torch.empty((world * t2.shape[0],) + tuple(t2.shape[1:]), dtype=t2.dtype)
        dist.all_gather_into_tensor(gathered, t2)                                 # one shape on every rank
        # a later, shorter batch (the ragged tail of an epoch) pads to the same capacity; a larger one is refused loudly
        t3, _, _ = sync.pad_to_capacity(tok[:1, :3], msk[:1, :3], dh[:1, :3])
        assert t3.shape == (16, 3)
        try:
            sync.pad_to_capacity(torch.zeros(3, 8, 3, dtype=torch.int64), torch.zeros(3, 8, 3, dtype=torch.uint8),
                                 torch.zeros(3, 8, 4))
            raise AssertionError("a batch above the fixed capacity must raise")
        except RuntimeError:
            pass
        # (1b) LoRA flat exchange: rank 1 has no gradient for the second parameter (no decoder frame selected): both
        #      ranks still issue one all-reduce of the same size, and the missing gradient counts as zero
        pa, pb = torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.zeros(2, 3))
        flat_sync = dp.GradSynchronizer([pa, pb])
        pa.grad = torch.full((4,), float(rank + 1))
        if rank == 0:
            pb.grad = torch.full((2, 3), 4.0)
        flat_sync.finish()
        assert torch.allclose(pa.grad, torch.full((4,), 1.5)) and torch.allclose(pb.grad, torch.full((2, 3), 2.0))
        # (2) every rank sees the same number of batches even when the dataset does not divide evenly
        data = [{"input_tokens": torch.ones(3 + i, 33, dtype=torch.long),
                 "input_masks": torch.ones(3 + i, 33, dtype=torch.bool),
                 "target_audio_tokens": torch.ones(3 + i, 32, dtype=torch.long)} for i in range(7)]
        mine = sum(1 for _ in iterate_batches(data, 2, True, rank, world, seed=1))
        counts = [None] * world
        dist.all_gather_object(counts, mine)
        assert counts == [2, 2], counts                                           # 7 samples -> 3 per rank -> 2 batches
        q.put((rank, "ok"))
    except Exception:  # noqa: BLE001
        import traceback
        q.put((rank, traceback.format_exc()))
    finally:
        dist.destroy_process_group()


def test_ragged_batches_share_shapes_and_step_counts_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ragged, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_bucket_views_for_direct_wgrad_writes_single_process():
    """grad_view / packed_view / mark_written: the wgrad GEMMs of the hand-written backward write straight into the
    all-reduce buckets.  Projections of a fused GEMM that are adjacent (in GEMM order) in the bucket layout get ONE
    [sum(rows), cols] view; anything else falls back to None (the caller then copies)."""
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path.insert(0, os.path.join(root, "csm-train-pytorch_b200"))
    from csm.training import dp
    q, k, v = (torch.nn.Parameter(torch.zeros(r, 8)) for r in (16, 4, 4))
    o, n = torch.nn.Parameter(torch.zeros(16, 8)), torch.nn.Parameter(torch.zeros(8))
    sync = dp.GradSynchronizer([q, k, v, o, n], bucket_bytes=1 << 20, force_buckets=True, bucket_order=[n, q, k, v, o])
    pv = sync.packed_view([q, k, v])
    assert pv is not None and pv.shape == (24, 8)
    assert sync.packed_view([k, q, v]) is None and sync.packed_view([q, k, v, n]) is None    # order / shape mismatch
    pv.copy_(torch.arange(24 * 8, dtype=torch.float32).view(24, 8))        # "the GEMM wrote the fused weight gradient"
    assert not sync.has_grad(q)
    for p in (q, k, v):
        sync.mark_written(p)
    assert sync.has_grad(q) and q.grad.data_ptr() == pv.data_ptr() and torch.equal(k.grad, pv[16:20])
    assert torch.equal(v.grad, pv[20:24])
    gv = sync.grad_view(o)
    gv.fill_(2.0)
    sync.mark_written(o)
    n.grad = torch.ones(8)                     # arrives through the autograd hook path in real runs; here: deliver
    sync.deliver(n, torch.ones(8))
    sync.finish()                              # world 1: nothing to reduce, state resets
    assert torch.equal(o.grad, torch.full((16, 8), 2.0)) and not sync._arrived
    # a second bucket boundary between w1 and w3 breaks adjacency -> None
    w1, w3 = torch.nn.Parameter(torch.zeros(32, 8)), torch.nn.Parameter(torch.zeros(32, 8))
    s2 = dp.GradSynchronizer([w1, w3], bucket_bytes=32 * 8 * 4, force_buckets=True, bucket_order=[w1, w3])
    assert s2.packed_view([w1, w3]) is None and s2.grad_view(w1) is not None

"""CPU suite, part 4: the N>1 path.  Two gloo ranks on 127.0.0.1, dialogues sharded by index, one
bucketed SUM allreduce overlapped with backward (ddp.GradAllReducer); the result must equal the
single-process step on the concatenated batch, and never-used parameters must keep
grad None (F8).  The CUDA library is replaced by the plain-C oracle backend in the workers."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
T, N, WORLD = 4, 4, 2


def _full_batch():
    g = torch.Generator().manual_seed(9)
    return torch.randn(T, N, 200, generator=g), torch.randint(0, 7, (T, N), generator=g)


def _worker(rank, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import oracle_backend
    from helpers import masked_ce, seeded_model
    from importlib import import_module
    import lsthm_b200
    ddp = import_module(lsthm_b200.__name__ + ".ddp")
    oracle_backend.install_plain()
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=WORLD)
    try:
        model = seeded_model("AT", 21).eval()
        x, lab = _full_batch()
        sh = slice(rank * N // WORLD, (rank + 1) * N // WORLD)
        # buckets in the order autograd finishes the gradients (observed on a dry step), so they fill front to back
        order = ddp.observe_grad_order(
            model, lambda: masked_ce(model(x[:, sh].contiguous()), lab[:, sh].reshape(-1), T, N // WORLD).backward())
        assert all(p.grad is None for p in model.parameters()) and len(order) == len(set(order)) > 10
        reducer = ddp.GradAllReducer(model, WORLD, bucket_bytes=256 << 10, order=order)
        for _ in range(2):                       # twice: zero_grad must re-arm the buckets
            reducer.zero_grad()
            probs = model(x[:, sh].contiguous())
            n_sh = N // WORLD
            loss = masked_ce(probs, lab[:, sh].reshape(-1), T, n_sh) * (n_sh / N)   # n_shard / n_global
            loss.backward()
            reducer.finish()
        if rank == 0:
            torch.save({n: (None if p.grad is None else p.grad.clone()) for n, p in model.named_parameters()}, out)
            assert len(reducer.buckets) >= 2
            # every bucket completed during the backward, in index order, and the last-finished gradient sits in the last bucket
            assert reducer.fire_order == list(range(len(reducer.buckets))), reducer.fire_order
            last = dict(model.named_parameters())[order[-1]]
            assert any(q is last for q in reducer._members[-1])
    finally:
        dist.destroy_process_group()


def test_two_rank_allreduce_equals_single_process(tmp_path, monkeypatch):
    import oracle_backend
    from helpers import e_inf, masked_ce, seeded_model
    out = str(tmp_path / "grads.pt")
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(port, out), nprocs=WORLD, join=True)
    sharded = torch.load(out)
    oracle_backend.install(monkeypatch)
    model = seeded_model("AT", 21).eval()
    x, lab = _full_batch()
    masked_ce(model(x), lab.reshape(-1), T, N).backward()
    for n, p in model.named_parameters():
        if p.grad is None:
            assert sharded[n] is None, n
        else:
            assert e_inf(sharded[n], p.grad) < 1e-5, n

"""GPU suite, multi-GPU part (skipped on a 1-GPU box): two NCCL ranks, dialogues sharded by index, the
bucketed gradient allreduce overlapped with backward must reproduce the single-process step on the
concatenated batch — through the real CUDA kernels this time."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))
T, N, WORLD = 12, 16, 2


def _batch():
    g = torch.Generator().manual_seed(4)
    return torch.randn(T, N, 712, generator=g), torch.randint(0, 6, (T, N), generator=g)


def _worker(rank, port, out):
    sys.path.insert(0, HERE)
    sys.path.insert(0, os.path.dirname(HERE))
    import torch.distributed as dist
    from importlib import import_module
    import lsthm_b200
    from helpers import masked_ce, seeded_model
    ddp = import_module(lsthm_b200.__name__ + ".ddp")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=WORLD, device_id=torch.device("cuda", rank))
    try:
        model = seeded_model("ATV", 31, f"cuda:{rank}").eval()
        reducer = ddp.GradAllReducer(model, WORLD, bucket_bytes=1 << 20)
        x, lab = _batch()
        sh = slice(rank * N // WORLD, (rank + 1) * N // WORLD)
        for _ in range(2):
            reducer.zero_grad()
            probs = model(x[:, sh].contiguous().cuda())
            loss = masked_ce(probs, lab[:, sh].reshape(-1).cuda(), T, N // WORLD) * (1.0 / WORLD)
            loss.backward()
            reducer.finish()
        torch.cuda.synchronize()
        if rank == 0:
            torch.save({n: (None if p.grad is None else p.grad.cpu()) for n, p in model.named_parameters()}, out)
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    from helpers import e_inf, masked_ce, seeded_model
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(29600 + os.getpid() % 2000, out), nprocs=WORLD, join=True)
    sharded = torch.load(out)
    model = seeded_model("ATV", 31, "cuda:0").eval()
    x, lab = _batch()
    masked_ce(model(x.cuda()), lab.reshape(-1).cuda(), T, N).backward()
    for n, p in model.named_parameters():
        if p.grad is None:
            assert sharded[n] is None, n
        else:
            assert e_inf(sharded[n], p.grad.cpu()) < 2e-4, n

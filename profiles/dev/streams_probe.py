"""Per-step device times of the ATV train step with the encoders on one stream / three streams, and the number of
cudaMalloc calls the caching allocator made during the timed steps (GPU box): python profiles/dev/streams_probe.py"""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import lsthm_b200
import bench

T, B = 110, 1024
dev = torch.device("cuda", 0)
torch.manual_seed(111)
model = lsthm_b200.HybridRNN_ATV.MARN().to(dev).train()
loss_fn = lsthm_b200.MaskedLoss(torch.nn.CrossEntropyLoss)
batches = [tuple(t.to(dev) for t in bench.synthetic_batch(111 + i, T, B, pinned=False, model="ATV")) for i in range(2)]


def step(i):
    b = batches[i & 1]
    model.zero_grad(set_to_none=True)
    loss_fn(model(b[0]), b[1], b[2]).backward()


for mode in sys.argv[1:] or ["0", "1", "0", "1"]:
    model.concurrent_encoders = mode == "1"
    for i in range(3):
        step(i)
    torch.cuda.synchronize()
    st0 = torch.cuda.memory_stats()
    evs = [torch.cuda.Event(enable_timing=True) for _ in range(31)]
    import time
    w0 = time.perf_counter()
    evs[0].record()
    for i in range(30):
        step(i)
        evs[i + 1].record()
    host_ms = (time.perf_counter() - w0) * 1e3 / 30
    torch.cuda.synchronize()
    print(f"host issue {host_ms:.2f} ms/step", end="; ")
    st1 = torch.cuda.memory_stats()
    ms = [evs[i].elapsed_time(evs[i + 1]) for i in range(30)]
    print(f"streams={mode} mean {sum(ms) / 30:.3f} min {min(ms):.3f} max {max(ms):.3f} ms;  cudaMalloc calls during the timed steps: "
          f"{st1['num_device_alloc'] - st0['num_device_alloc']}, frees {st1['num_device_free'] - st0['num_device_free']}, "
          f"reserved {st1['reserved_bytes.all.current'] / 2**30:.2f} GiB", flush=True)
    print("   ", " ".join(f"{m:.2f}" for m in ms), flush=True)

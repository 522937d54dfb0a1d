"""Upper bound of what a CUDA-graph replay of the whole ATV train step buys over eager launches (GPU box):
python profiles/dev/graph_probe.py.  The captured step replays the dropout seeds of the capture (host constants in the
descriptors), so this is a timing probe only."""
import os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import lsthm_b200
import bench

T, B = 110, 1024
dev = torch.device("cuda", 0)
torch.manual_seed(111)
model = lsthm_b200.HybridRNN_ATV.MARN().to(dev).train()
loss_fn = lsthm_b200.MaskedLoss(torch.nn.CrossEntropyLoss)
batch = tuple(t.to(dev) for t in bench.synthetic_batch(111, T, B, pinned=False, model="ATV"))


def step():
    model.zero_grad(set_to_none=True)
    loss = loss_fn(model(batch[0]), batch[1], batch[2])
    loss.backward()
    return loss


def timed(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    w0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    w_issue = time.perf_counter() - w0
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, w_issue * 1e3 / n


for _ in range(3):
    step()
print("eager: %.3f ms/step on the device, host issue time %.3f ms/step" % timed(step, 10), flush=True)
s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(3):
        step()
torch.cuda.current_stream().wait_stream(s)
torch.cuda.synchronize()
g = torch.cuda.CUDAGraph()
model.zero_grad(set_to_none=True)
with torch.cuda.graph(g):
    static_loss = loss_fn(model(batch[0]), batch[1], batch[2])
    static_loss.backward()
g.replay()
torch.cuda.synchronize()
print("graph loss", float(static_loss), "eager loss", float(step()), flush=True)
print("graph: %.3f ms/step on the device, host issue time %.3f ms/step" % timed(g.replay, 10), flush=True)
gn = sum(float(p.grad.abs().sum()) for p in model.parameters() if p.grad is not None)
print("grad l1 after replay", gn)

"""Where the HOST time of one eager ATV train step goes (cProfile over 10 steps; GPU box): python profiles/dev/host_profile.py"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
import torch
import lsthm_b200
import bench

T, B = 110, int(os.environ.get("B", "1024"))
dev = torch.device("cuda", 0)
torch.manual_seed(111)
model = lsthm_b200.HybridRNN_ATV.MARN().to(dev).train()
loss_fn = lsthm_b200.MaskedLoss(torch.nn.CrossEntropyLoss)
b = tuple(t.to(dev) for t in bench.synthetic_batch(111, T, B, pinned=False, model="ATV"))


def step():
    model.zero_grad(set_to_none=True)
    loss_fn(model(b[0]), b[1], b[2]).backward()


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(10):
    step()
host = (time.perf_counter() - t0) / 10
torch.cuda.synchronize()
print(f"host issue {host * 1e3:.2f} ms/step (device-bound total {(time.perf_counter() - t0) / 10 * 1e3:.2f})")
pr = cProfile.Profile()
pr.enable()
for _ in range(10):
    step()
pr.disable()
torch.cuda.synchronize()
st = pstats.Stats(pr)
st.sort_stats("tottime").print_stats(35)

#!/usr/bin/env python
"""bench.py — utterances/sec of one fwd+bwd of the drop-in HybridRNN_ATV module (BASELINE.json
configs[1]: audio+text+visual, cross-modal attention + fusion, fp32, 1xB200; configs[4] at N>1).

    python bench.py [--gpus N] [--steps K] [--warmup W]                  # our arm (CUDA, C ABI)
    python bench.py --impl reference [--steps K] [--warmup W]           # the UNMODIFIED reference on the host CPU (oracle/_ref)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

A "step" = zero grads -> forward -> MaskedLoss(CrossEntropy) -> backward (-> gradient allreduce
complete on every rank when N>1) on one synthetic IEMOCAP-shaped batch x[110, 1024, 712] per GPU
(weak scaling), train mode.  Prints ONE JSON line (rank 0).  At N = 1 the line also carries, as extra
fields, the other arms SURVEY.md §8d asks for: the reference run eagerly on the same B200, config 1
(HybridRNN_AT, B = 32, CPU), config 4 (DialogueRNN BiModel: CPU and eager B200), config 3 (MARN1_sps,
fp32 and bf16) and the optimizer step timed separately.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "utterances/sec fwd+bwd HybridRNN_ATV"
UNIT = "utterances/s"
T_LEN, BATCH, D_IN, N_CLS = 110, 1024, 712, 6
# --model sps (BASELINE.json configs[2] shapes, fp32): per-direction FLOPs of the fused cell kernels
SPS_FLOP_FWD = 524_288 + 786_432 + 82_432
SPS_FLOP_BWD = 524_288 + 786_432 + 3 * 82_432
# algorithmic FLOPs per utterance of the serial chain: fwd 999,936; the BPTT adjoint
# chain is the five transposed products of the same weights = the same count.
FLOP_FWD_PER_UTT = 999_936
FLOP_BWD_PER_UTT = 999_936


SPEAKER_KINDS = ("sps", "onlysp", "nsps", "no_en")
# GRU speaker-state cell (lsthm_onlysp / lsthm_nsps), per direction: W_hh product + two LSTHM1 gate products + collapsed attention
GSP_FLOP_FWD = 98_304 + 786_432 + 82_432
GSP_FLOP_BWD = 98_304 + 786_432 + 3 * 82_432


def model_name(kind):
    return {"ATV": "HybridRNN_ATV", "sps": "MARN1_sps", "onlysp": "MARN1_onlysp", "nsps": "MARN1_nsps", "no_en": "MARN1_no_en"}[kind]


def metric_name(kind):
    return f"utterances/sec fwd+bwd {model_name(kind)}"


def workload_config(kind, T, B, world, info=None):
    """The `config` object of the JSON line — the same for our arm and for the reference arm (which times a bounded sample of
    this workload on the host cores and says so in `cpu_baseline.sample`)."""
    din = D_IN if kind == "ATV" else 1124
    cfg = {"workload": f"{model_name(kind)} fwd+bwd (train mode), x[{T},{B},{din}] fp32 per GPU, uniform L={T}, "
                       f"MaskedLoss(CrossEntropy); inputs {T * B * din * 4 / 1e6:.0f} MB per step > 126 MB L2, "
                       f"two alternating batches", "per_gpu_batch": B, "seq_len": T, "parallelism": f"dp{world}"}
    if info is not None:
        cfg.update({"grid": info["grid"], "block": info["block"], "rows_per_cta": info["rows"], "group": info.get("group")})
    return cfg


def launch_info(kind, T, B):
    """Launch geometry of the dominant kernel pair (host-side query of the library; no GPU needed)."""
    import lsthm_b200
    from importlib import import_module
    _l = import_module(lsthm_b200.__name__ + "._lib")
    if kind == "ATV":
        info = _l.mab_launch_info(_l.make_desc(T, B, (128, 16, 64), (16, 128, 100)))
        info["rows"] = info["dialogues_per_group"]
    elif kind in ("onlysp", "nsps", "no_en"):
        info = _l.gsp_launch_info(_l.make_gsp_desc(T, B, 0 if kind == "onlysp" else 1))
    else:
        info = _l.sps_launch_info(_l.make_sps_desc(T, B))
    return info


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm_gbs=d["hbm_gbs"], bf16_burst=d["bf16_tflops"], bf16_sustained=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    sm_max_mhz=d.get("sm_max_mhz", 1965.0), source="measured")
    return dict(hbm_gbs=6650.0, bf16_burst=1590.0, bf16_sustained=1400.0, sm_max_mhz=1965.0, source="fallback")


class ClockSampler:
    """Samples SM clocks / throttle reasons with nvidia-smi during the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                                          "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), line.strip()))

    def stop(self, t0, t1):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for ts, line in self.rows:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 6 or not (t0 - 0.05 <= ts <= t1 + 0.15):
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for n, v in zip(names, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def synthetic_batch(seed, T, B, device=None, pinned=False, model="ATV"):
    """Seeded IEMOCAP-shaped batch (uniform set: every dialogue L = 110).  For the sps
    model also a two-speaker one-hot qmask (first speaker Bernoulli(0.5), switch probability 0.6)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.randn(T, B, D_IN if model == "ATV" else 1124, generator=g)
    p = torch.tensor([144, 245, 384, 170, 299, 381], dtype=torch.float32) / 1623.0
    labels = torch.multinomial(p, T * B, replacement=True, generator=g)
    umask = torch.ones(B, T)
    out = [x, labels, umask]
    if model in SPEAKER_KINDS:
        spk = torch.randint(0, 2, (B,), generator=g)
        qmask = torch.zeros(T, B, 2)
        for t in range(T):
            spk = torch.where(torch.rand(B, generator=g) < 0.6, 1 - spk, spk)
            qmask[t, torch.arange(B), spk] = 1
        out.append(qmask)
    if pinned:
        out = [t.pin_memory() for t in out]
    if device is not None:
        out = [t.to(device) for t in out]
    return tuple(out)


# ------------------------------------------------------------------------------------------------
# reference arm / cpu_baseline / eager-GPU arms: the UNMODIFIED reference modules.  `make -C oracle ref` (run by
# __graft_entry__.build() where /root/reference exists) stages byte-for-byte copies of the hot-path files into the
# git-ignored oracle/_ref/, which travels to the GPU box; oracle/ref_shim.py repairs the import paths in sys.modules.
# This is the one place bench.py executes oracle/ code: as the thing compared against, never on the product path.
# ------------------------------------------------------------------------------------------------
def reference_ns():
    from oracle import ref_shim
    if not ref_shim.reference_available():
        raise RuntimeError("the reference is neither at /root/reference nor staged in oracle/_ref (run `make -C oracle ref`)")
    return ref_shim.load_reference()


def reference_step_fn(kind, B, device="cpu", T=T_LEN, seed=111):
    """(step(), utterances per step) for one fwd + MaskedLoss + backward of the reference class `kind`, train mode (dropout on),
    default init under seed 111, on the same synthetic batch generator as our arm."""
    ns = reference_ns()
    torch.manual_seed(seed)
    loss_fn = ns.MaskedLoss(torch.nn.CrossEntropyLoss)
    if kind in ("ATV", "AT"):
        model = (ns.MARN_ATV if kind == "ATV" else ns.MARN_AT)().to(device).train()
        g = torch.Generator().manual_seed(seed)
        x = torch.randn(T, B, D_IN if kind == "ATV" else 200, generator=g).to(device)
        labels = torch.randint(0, 6, (T * B,), generator=g).to(device)
        umask = torch.ones(B, T, device=device)

        def step():
            model.zero_grad(set_to_none=True)
            loss = loss_fn(model(x), labels, umask)       # HybridRNN_AT(V).py: forward(x) -> [T*B, C] probabilities, time-major
            loss.backward()
            return loss
    elif kind in SPEAKER_KINDS:
        model = {"sps": lambda: ns.MARN1_sps(6), "onlysp": lambda: ns.MARN1_onlysp(6), "nsps": lambda: ns.MARN1_nsps(6, "IEMOCAP"),
                 "no_en": lambda: ns.MARN1_no_en(6, "IEMOCAP")}[kind]()
        model = model.to(device).train()
        x, labels, umask, qmask = [t.to(device) for t in synthetic_batch(seed, T, B, model="sps")]
        lab_bm = labels.view(T, B).t().reshape(-1)        # lsthm_sps.py:392-393 returns batch-major rows

        def step():
            model.zero_grad(set_to_none=True)
            loss = loss_fn(model(x, qmask, umask)[0], lab_bm, umask)
            loss.backward()
            return loss
    elif kind == "DialogueRNN":
        # ctor arguments of model_trainer.py:35-47; call convention of model_trainer_d.py:63-67 (att2=True)
        model = ns.BiModel(712, 500, 500, 300, 300, n_classes=6, listener_state=True, context_attention="general",
                           dropout_rec=0.1, dropout=0.1).to(device).train()
        x, labels, umask, qmask = [t.to(device) for t in synthetic_batch(seed, T, B, model="sps")]
        x = x[:, :, :712].contiguous()
        lab_bm = labels.view(T, B).t().reshape(-1)

        def step():
            model.zero_grad(set_to_none=True)
            log_prob = model(x, qmask, umask, att2=True)[0]
            lp = log_prob.transpose(0, 1).contiguous().view(-1, log_prob.size(2))
            loss = loss_fn(lp, lab_bm, umask)
            loss.backward()
            return loss
    else:
        raise ValueError(kind)
    return step, T * B


def best_threads(step, candidates):
    best = None
    for n in candidates:
        torch.set_num_threads(n)
        step()
        t = time.perf_counter(); step(); dt = time.perf_counter() - t
        if best is None or dt < best[1]:
            best = (n, dt)
    torch.set_num_threads(best[0])
    return best[0]


def time_cpu(kind, B, steps, sweep=True):
    """Reference `kind` on the host cores: thread sweep, then `steps` timed steps; returns a cpu_baseline-shaped dict."""
    ncpu = os.cpu_count() or 1
    step, utt = reference_step_fn(kind, B, "cpu")
    cands = sorted({1, min(4, ncpu), min(8, ncpu), ncpu}) if sweep else [min(8, ncpu)]
    if sweep:
        n = best_threads(step, cands)
    else:
        n = cands[0]
        torch.set_num_threads(n)
        step()
    ts = []
    for _ in range(steps):
        t = time.perf_counter(); step(); ts.append(time.perf_counter() - t)
    from oracle import ref_shim
    return {"value": utt / min(ts), "unit": UNIT, "cores": n, "kind": "reference", "source": "oracle/_ref" if ref_shim.REF_KIND == "staged" else ref_shim.REF_ROOT,
            "ms_per_step": 1e3 * min(ts),
            "sample": f"unmodified reference {kind} (train mode, dropout on) fwd + MaskedLoss + bwd, x[{T_LEN},{B},*] fp32, best of {steps} "
                      f"after thread sweep {cands} on {ncpu} host cpus"}


def time_eager_gpu(kind, B, steps, warm, dev):
    """The same unmodified reference module run by PyTorch eager ON THE B200 (the existing GPU implementation)."""
    step, utt = reference_step_fn(kind, B, dev)
    for _ in range(warm):
        step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    return {"value": utt / (ms * 1e-3), "unit": UNIT, "ms_per_step": ms, "batch": B, "steps": steps,
            "what": f"unmodified reference {kind} (oracle/_ref), PyTorch eager on the same B200, train mode, x[{T_LEN},{B},*] fp32"}


def run_cpu_baseline(steps=2, B=32, model_kind="ATV"):
    if model_kind in SPEAKER_KINDS:
        B = 8                                  # these reference models loop over the batch in Python: ~6x slower per utterance
    return time_cpu(model_kind, B, steps)


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    B = 32 if args.model == "ATV" else 8
    ncpu = os.cpu_count() or 1
    step, utt = reference_step_fn(args.model, B, "cpu")
    cands = sorted({1, min(4, ncpu), min(8, ncpu), min(16, ncpu), ncpu})
    n = best_threads(step, cands)
    for _ in range(max(0, args.warmup - len(cands))):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = utt * args.steps / dt
    from oracle import ref_shim
    src = "oracle/_ref" if ref_shim.REF_KIND == "staged" else ref_shim.REF_ROOT
    try:
        cfg = workload_config(args.model, args.seq, args.batch, args.gpus, launch_info(args.model, args.seq, args.batch))
    except Exception:
        cfg = workload_config(args.model, args.seq, args.batch, args.gpus)
    sample = (f"the unmodified reference modules ({src}) on the host cores, train mode, {args.model} x[{T_LEN},{B},*] fp32 per step "
              f"(a bounded sample of the {BATCH}-dialogue workload), {n} torch threads (best of sweep {cands}; {ncpu} host cpus)")
    print(json.dumps({
        "impl": "reference", "metric": metric_name(args.model), "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": n, "kind": "reference", "source": src, "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch.distributed as dist
    import lsthm_b200
    from importlib import import_module
    rec = import_module(lsthm_b200.__name__ + ".recurrence")
    ddp = import_module(lsthm_b200.__name__ + ".ddp")
    mm3 = import_module(lsthm_b200.__name__ + ".mm3")
    fat = import_module(lsthm_b200.__name__ + ".fused_attention")
    fdl = import_module(lsthm_b200.__name__ + ".fused_dln")
    lss = import_module(lsthm_b200.__name__ + ".loss")
    sqa = import_module(lsthm_b200.__name__ + ".seq_attention")

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node N for --gpus N > 1")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # the contract is ONE line on stdout: NCCL prints its "NCCL version ..." banner to stdout at NCCL_DEBUG=VERSION
        # (what this image sets) and honours NCCL_DEBUG_FILE only above that level -> raise to WARN and log to stderr
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)
    T, B = args.seq, args.batch
    kind = args.model
    _lib = import_module(lsthm_b200.__name__ + "._lib")
    _lib.set_precision("bf16" if args.dtype == "bf16" else "fp32")

    def make_model(k):
        torch.manual_seed(111)
        if k == "ATV":
            return lsthm_b200.HybridRNN_ATV.MARN().to(dev).train()
        m = {"sps": lambda: lsthm_b200.lsthm_sps.MARN1_sps(6), "onlysp": lambda: lsthm_b200.lsthm_onlysp.MARN1_onlysp(6),
             "nsps": lambda: lsthm_b200.lsthm_nsps.MARN1_nsps(6, "IEMOCAP"),
             "no_en": lambda: lsthm_b200.lsthm_no_en.MARN1_no_en(6, "IEMOCAP")}[k]()
        # the stock init sets every attention projection / fusion scalar to ones (lsthm_sps.py:52-54,82-84,340-346):
        # degenerate softmaxes; perturb them as the parity tests do so the timed arithmetic is representative
        gpert = torch.Generator().manual_seed(114)
        with torch.no_grad():
            for prm in m.parameters():
                if bool((prm == 1).all()):
                    prm.add_(0.1 * torch.randn(prm.shape, generator=gpert))
        return m.to(dev).train()
    model = make_model(kind)
    loss_fn = lsthm_b200.MaskedLoss(torch.nn.CrossEntropyLoss)

    def forward_of(m, k, batch):
        if k == "ATV":
            return m(batch[0])
        return m(batch[0], batch[3], batch[2])[0]

    def forward(batch):
        return forward_of(model, kind, batch)
    # two host batches (pinned) so consecutive steps see different data; device-resident copies for `value`
    host = [synthetic_batch(111 + 7 * rank + i, T, B, pinned=True, model="ATV" if kind == "ATV" else "sps") for i in range(2)]
    resident = [tuple(t.to(dev) for t in hb) for hb in host]
    reducer = None
    if world > 1:
        # gradient buckets in the order autograd finishes them (one dry step).  Default = the measured best on 8 x B200
        # (profiles/r02/scale_ab_n8.log): ONE bucket, reduced after the backward (14.81 ms per step; 2 MB buckets overlapped
        # with the backward: 15.22 ms; one GPU: 14.65 ms).  The two environment variables are experiment knobs of this script.
        order = ddp.observe_grad_order(model, lambda: loss_fn(forward(resident[0]), resident[0][1], resident[0][2]).backward())
        bb = int(os.environ.get("LSTHM_DDP_BUCKET_MB", "64")) << 20
        reducer = ddp.GradAllReducer(model, world, bucket_bytes=bb, order=order, overlap=os.environ.get("LSTHM_DDP_OVERLAP", "0") == "1")
    utt_per_step = T * B * world

    def step_resident(i):
        batch = resident[i & 1]
        labels, umask = batch[1], batch[2]
        if reducer is not None:
            reducer.zero_grad()
        else:
            model.zero_grad(set_to_none=True)
        probs = forward(batch)
        loss = loss_fn(probs, labels, umask)
        if reducer is not None:
            loss = loss * (1.0 / world)
        loss.backward()
        if reducer is not None:
            reducer.finish()
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing (value) with per-kernel CUDA events on the launching stream ----
    for i in range(args.warmup):
        step_resident(i)
    barrier()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start(); time.sleep(0.25)
    rec.kernel_events = {"fwd": [], "bwd": []}
    for cnt in (rec.launch_counter, mm3.launches, fat.launches, fdl.launches, lss.launches, sqa.launches_x):
        for k in cnt:
            cnt[k] = 0
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    wall0 = time.time()
    e0.record()
    for i in range(args.steps):
        step_resident(i)
    e1.record()
    barrier()
    wall1 = time.time()
    ms = max_over_ranks(e0.elapsed_time(e1))
    launches = (sum(rec.launch_counter.values()) + sum(mm3.launches.values()) + sum(fat.launches.values())
                + sum(fdl.launches.values()) + sum(lss.launches.values()) + sum(sqa.launches_x.values())) * world
    kev, rec.kernel_events = rec.kernel_events, None
    # the ~60 small tensor-core launches of a step are event-timed in two EXTRA steps outside the timed region
    # (an event pair per launch would perturb `value`)
    # ... and on the SINGLE-stream schedule: an event pair around a launch measures its stream's elapsed time, which with the
    # model's branches on concurrent streams includes the other branches' kernels sharing the SMs
    TC_STEPS = 2
    concurrent = getattr(model, "concurrent_encoders", None)
    if concurrent:
        model.concurrent_encoders = False
    mm3.events, fat.events = [], []
    for i in range(TC_STEPS):
        step_resident(i)
    torch.cuda.synchronize()
    gev, mm3.events = mm3.events, None
    aev, fat.events = fat.events, None
    if concurrent:
        model.concurrent_encoders = True

    def tc_summary(evs):
        """Tensor-core kernels of the step: time, fp32-equivalent rate, and rate of the bf16 UMMAs actually
        issued (3 per product term: hi.hi + hi.lo + lo.hi) against the measured bf16 peak."""
        tot_ms = sum(a.elapsed_time(b) for a, b, _ in evs)
        fl = sum(f for _, _, f in evs)
        if tot_ms <= 0:
            return None
        tf = fl / (tot_ms * 1e-3) / 1e12
        return {"launches_per_step": len(evs) / TC_STEPS, "ms_per_step": tot_ms / TC_STEPS, "fp32_equiv_tflops": tf,
                "bf16_umma_tflops": 3 * tf, "frac_of_bf16_peak": 3 * tf / peaks()["bf16_sustained"]}
    kms = {k: (sum(a.elapsed_time(b) for a, b in v) / max(1, len(v))) for k, v in kev.items()}
    clocks = sampler.stop(wall0, wall1) if sampler else None
    value = utt_per_step * args.steps / (ms * 1e-3)

    e2e_val = e2e_ms = None
    h2d = d2h = 0
    if not args.no_e2e:
        # ---- end to end through the public module API: pinned host inputs -> H2D -> step -> loss D2H ----
        copy_stream = torch.cuda.Stream(device=dev)
        dbuf = [tuple(torch.empty_like(t, device=dev) for t in host[0]) for _ in range(2)]
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]

        def prefetch(i):
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[i & 1])
                for dst, src in zip(dbuf[i & 1], host[i & 1]):
                    dst.copy_(src, non_blocking=True)
                ready[i & 1].record(copy_stream)

        def step_e2e(i, last):
            cur = torch.cuda.current_stream()
            cur.wait_event(ready[i & 1])
            if not last:
                prefetch(i + 1)               # next step's H2D overlaps this step's compute
            batch = dbuf[i & 1]
            labels, umask = batch[1], batch[2]
            if reducer is not None:
                reducer.zero_grad()
            else:
                model.zero_grad(set_to_none=True)
            loss = loss_fn(forward(batch), labels, umask)
            if reducer is not None:
                loss = loss * (1.0 / world)
            loss.backward()
            if reducer is not None:
                reducer.finish()
            freed[i & 1].record(cur)
            # D2H read of the step's result, every step: asynchronous copy of the loss into pinned memory behind the step,
            # consumed by the host ONE step later (while the next step is already queued), as a training loop that logs its
            # loss does; the last step's value is read before the clock stops.  A blocking .item() here idles the GPU for the
            # ~0.8 ms the host needs to get the next step's first launches out (measured: 16.05 vs 15.26 ms per step).
            loss_host[i & 1].copy_(loss.detach().reshape(1), non_blocking=True)
            loss_ready[i & 1].record(cur)
            if i > 0:
                loss_ready[(i - 1) & 1].synchronize()
                losses.append(float(loss_host[(i - 1) & 1][0]))
            if last:
                loss_ready[i & 1].synchronize()
                losses.append(float(loss_host[i & 1][0]))

        loss_host = [torch.empty(1, dtype=torch.float32).pin_memory() for _ in range(2)]
        loss_ready = [torch.cuda.Event() for _ in range(2)]
        losses = []
        for ev in freed:
            ev.record(torch.cuda.current_stream())
        n_e2e_warm = max(2, min(args.warmup, 3))
        prefetch(0)
        for i in range(n_e2e_warm):
            step_e2e(i, False)
        barrier()
        n_before = len(losses)
        t0 = time.perf_counter()
        base = n_e2e_warm
        for i in range(args.steps):
            step_e2e(base + i, i == args.steps - 1)
        barrier()
        e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3)
        # every timed step's loss reached the host inside the timed region (+ the last warm-up step's, read one step late)
        assert len(losses) - n_before == args.steps + 1 and all(l == l for l in losses), (len(losses), n_before)
        e2e_val = utt_per_step * args.steps / (e2e_ms * 1e-3)
        h2d = sum(t.numel() * t.element_size() for t in host[0]) * world
        d2h = 4 * world

    # ---- N > 1: guarded self-check that the sharded step equals the single-process step (the 2-GPU pytest is skipped on the
    #      driver's 1-GPU test box, so the multi-GPU runs carry the evidence): same seeded small batch on every rank, rank r
    #      takes its dialogue shard through GradAllReducer, and compares the reduced gradients with the full batch's ----
    extras = {}
    if world > 1:
        try:
            Tc, Nc = 12, 8 * world
            gc_ = torch.Generator().manual_seed(4)
            xc_, lc_ = torch.randn(Tc, Nc, D_IN, generator=gc_).to(dev), torch.randint(0, 6, (Tc, Nc), generator=gc_).to(dev)
            torch.manual_seed(31)
            mref = lsthm_b200.HybridRNN_ATV.MARN().to(dev).eval()
            torch.manual_seed(31)
            mshd = lsthm_b200.HybridRNN_ATV.MARN().to(dev).eval()
            um = torch.ones(Nc, Tc, device=dev)
            loss_fn(mref(xc_), lc_.reshape(-1), um).backward()
            red2 = ddp.GradAllReducer(mshd, world, bucket_bytes=1 << 20)
            sh = slice(rank * 8, rank * 8 + 8)
            red2.zero_grad()
            (loss_fn(mshd(xc_[:, sh].contiguous()), lc_[:, sh].reshape(-1), um[sh]) * (1.0 / world)).backward()
            red2.finish()
            worst = 0.0
            for (n1, p1), (_, p2) in zip(mref.named_parameters(), mshd.named_parameters()):
                if p1.grad is None:
                    assert p2.grad is None, n1
                    continue
                worst = max(worst, float((p1.grad - p2.grad).abs().max() / p1.grad.abs().max().clamp_min(1e-30)))
            wt = torch.tensor([worst], device=dev)
            dist.all_reduce(wt, op=dist.ReduceOp.MAX)
            extras["ddp_selfcheck"] = {"max_rel_grad_err": float(wt.item()), "ok": bool(wt.item() < 2e-4), "world": world,
                                       "what": f"ATV x[{Tc},{Nc},712] eval: NCCL-sharded step (8 dialogues per rank, bucketed allreduce) vs the "
                                               "single-process step on the whole batch, scale-relative max over all parameter gradients"}
            del mref, mshd, red2
        except Exception as e:
            extras["ddp_selfcheck"] = {"ok": False, "error": repr(e)[:300]}

    # ---- extra arms (N = 1 only; everything the timed region above does not include) ----
    if world == 1 and not args.no_extras:
        # (1) optimizer step, timed separately: fused Adam on the flat gradient buckets of a fresh copy
        try:
            m2 = make_model(kind)
            red = ddp.GradAllReducer(m2, 1, flatten_params=True)
            opt = ddp.FusedAdam(red, lr=1e-3, weight_decay=2e-5)
            bt = resident[0]
            loss_fn(forward_of(m2, kind, bt), bt[1], bt[2]).backward()
            red.finish()
            for _ in range(3):
                opt.step()
            torch.cuda.synchronize()
            o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            o0.record()
            for _ in range(20):
                opt.step()
            o1.record()
            torch.cuda.synchronize()
            extras["optimizer"] = {"ms_per_step": o0.elapsed_time(o1) / 20, "what": "FusedAdam (lr 1e-3, wd 2e-5; model_trainer.py:82) on the "
                                   f"flat buckets: {len(red.buckets)} launches per step, {sum(b.numel() for b in red.buckets)} parameters; NOT part of `value`"}
            del m2, red, opt
        except Exception as e:                                        # an extra arm must never cost the headline line
            extras["optimizer"] = {"error": repr(e)[:200]}
        # (2) config 3: MARN1_sps at its shapes, fp32 and bf16 (the headline model's numbers are above)
        if kind == "ATV":
            sps = {}
            for dt in ("f32", "bf16"):
                try:
                    _lib.set_precision("bf16" if dt == "bf16" else "fp32")
                    ms_model = make_model("sps")
                    sb = synthetic_batch(111, T, B, device=dev, model="sps")

                    def sps_step():
                        ms_model.zero_grad(set_to_none=True)
                        loss_fn(forward_of(ms_model, "sps", sb), sb[1].view(T, B).t().reshape(-1), sb[2]).backward()
                    for _ in range(3):
                        sps_step()
                    torch.cuda.synchronize()
                    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record()
                    for _ in range(5):
                        sps_step()
                    s1.record()
                    torch.cuda.synchronize()
                    sm = s0.elapsed_time(s1) / 5
                    sps[dt] = {"ms_per_step": sm, "value": T * B / (sm * 1e-3), "unit": UNIT}
                    del ms_model
                except Exception as e:
                    sps[dt] = {"error": repr(e)[:200]}
            _lib.set_precision("bf16" if args.dtype == "bf16" else "fp32")
            sps["workload"] = f"MARN1_sps(6) fwd+bwd (train mode), x[{T},{B},1124], qmask[{T},{B},2]; bf16 = bf16 operands in the time-parallel products, recurrence/softmax/LayerNorm fp32"
            extras["config3_sps"] = sps
            # (2b) the GRU speaker-state members of the family (lsthm_onlysp = train.py's default model, lsthm_nsps), fp32
            var = {}
            for vk in ("onlysp", "nsps", "no_en"):
                try:
                    vm = make_model(vk)
                    sb = synthetic_batch(111, T, B, device=dev, model="sps")

                    def var_step():
                        vm.zero_grad(set_to_none=True)
                        loss_fn(forward_of(vm, vk, sb), sb[1].view(T, B).t().reshape(-1), sb[2]).backward()
                    for _ in range(3):
                        var_step()
                    torch.cuda.synchronize()
                    rec.kernel_events = {"fwd": [], "bwd": []}
                    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    s0.record()
                    for _ in range(5):
                        var_step()
                    s1.record()
                    torch.cuda.synchronize()
                    sm = s0.elapsed_time(s1) / 5
                    kv = {k: sum(a.elapsed_time(b) for a, b in v) / max(len(v), 1) for k, v in rec.kernel_events.items()}
                    rec.kernel_events = None
                    var[vk] = {"ms_per_step": sm, "value": T * B / (sm * 1e-3), "unit": UNIT,
                               "cell_kernel_ms": {k: round(v, 4) for k, v in kv.items()}}
                    del vm
                except Exception as e:
                    var[vk] = {"error": repr(e)[:200]}
            var["workload"] = f"MARN1_onlysp(6) / MARN1_nsps(6) / MARN1_no_en(6) fwd+bwd (train mode), x[{T},{B},1124], qmask[{T},{B},2], fp32; cell_kernel_ms = per launch (one direction)"
            extras["gru_variants"] = var
        # (2c) the ragged set of SURVEY.md §8d (len = clip(round(N(52.4, 17.4)), 8, 110)): real utterances per second with the
        #      reference's random batching (every batch padded to its longest dialogue) vs the length-bucketed batches of
        #      pipeline.LengthBucketBatchSampler (f-4), on the headline model — 8 steps of B dialogues from one pool
        if kind == "ATV":
            try:
                pl = import_module(lsthm_b200.__name__ + ".pipeline")
                gl = torch.Generator().manual_seed(111)
                pool = 8 * B
                lens = torch.clamp(torch.round(52.4 + 17.4 * torch.randn(pool, generator=gl)), 8, 110).int().tolist()
                xs_full = torch.randn(T, B, D_IN, device=dev)
                lab_full = torch.randint(0, N_CLS, (T * B,), device=dev)
                rag = {}
                for label, pool_batches in (("random_batches", 1), ("length_bucketed", 8)):
                    bs = pl.LengthBucketBatchSampler(lens, B, pool_batches=pool_batches, seed=111)
                    batches = [b for b in bs if len(b) == B]
                    plans = []
                    for b in batches:
                        bl = torch.tensor([lens[i] for i in b])
                        Lb = int(bl.max())
                        um = (torch.arange(Lb)[None, :] < bl[:, None]).float().to(dev)          # [B, Lb]
                        xb = (xs_full[:Lb] * um.t().unsqueeze(-1)).contiguous()                  # zero on padding, as pad_sequence gives
                        plans.append((xb, lab_full[:Lb * B], um, int(bl.sum())))

                    def rag_epoch():
                        for xb, lb, um, _ in plans:
                            model.zero_grad(set_to_none=True)
                            loss_fn(model(xb), lb, um).backward()
                    rag_epoch()
                    torch.cuda.synchronize()
                    r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    r0.record()
                    rag_epoch()
                    r1.record()
                    torch.cuda.synchronize()
                    ms_ep = r0.elapsed_time(r1)
                    real = sum(p[3] for p in plans)
                    padded = sum(p[0].shape[0] * B for p in plans)
                    rag[label] = {"real_utterances_per_s": real / (ms_ep * 1e-3), "padded_positions_per_s": padded / (ms_ep * 1e-3),
                                  "padding_fraction": 1.0 - real / padded, "ms_per_step": ms_ep / len(plans), "steps": len(plans)}
                rag["workload"] = (f"HybridRNN_ATV fwd+bwd, {pool} synthetic dialogues with len = clip(round(N(52.4, 17.4)), 8, 110) in steps of {B}; "
                                   "each step padded to its longest dialogue (dataloader.py:45-47)")
                extras["ragged_set"] = rag
                del xs_full, plans
            except Exception as e:
                extras["ragged_set"] = {"error": repr(e)[:200]}
        # (3) the unmodified reference, PyTorch eager on this B200 and configs 1 and 4
        for name, fn in (("reference_eager_gpu", lambda: time_eager_gpu(kind, B, 3, 2, dev)),
                         ("config1_AT_cpu", lambda: time_cpu("AT", 32, 2)),
                         ("config4_dialoguernn_cpu", lambda: time_cpu("DialogueRNN", 32, 1, sweep=False)),
                         ("config4_dialoguernn_eager_gpu", lambda: time_eager_gpu("DialogueRNN", 32, 3, 1, dev))):
            try:
                extras[name] = fn()
            except Exception as e:
                extras[name] = {"error": repr(e)[:200]}
            torch.cuda.empty_cache()

    if rank == 0:
        pk = peaks()
        _l = import_module(lsthm_b200.__name__ + "._lib")
        dom = "bwd" if kms["bwd"] >= kms["fwd"] else "fwd"
        if kind == "ATV":
            flop_utt = FLOP_BWD_PER_UTT if dom == "bwd" else FLOP_FWD_PER_UTT
            kname = f"mab_{dom}_kernel"
            info = _l.mab_launch_info(_l.make_desc(T, B, (128, 16, 64), (16, 128, 100)))
            info["rows"] = info["dialogues_per_group"]
            D, G, R, MH = 208, 832, 244, 64
            # HBM bytes per utterance of one launch, fp32.  "algorithmic" = SURVEY.md §8(d)'s fused schedule (features once per
            # pass, only c,h,z stashed);  "scheduled" = what this kernel pair moves (it stashes gates, logits and per-head
            # products instead of recomputing them: DESIGN.md §3.1)
            by_alg = {"fwd": 2848 + 2496, "bwd": 2848 + 2496 + 1664 + 2848}
            by = {"fwd": 4 * (G + MH + D + MH + D + D + G + G + 8 + 4 * MH),
                  "bwd": 4 * (D + MH + MH + MH + 2 * D + D + G + G + 8 + 4 * MH + G + G + MH + G)}
            din = D_IN
            passes = 3                                              # hi.hi + hi.lo + lo.hi per product term
        elif kind in ("onlysp", "nsps", "no_en"):
            flop_utt = GSP_FLOP_BWD if dom == "bwd" else GSP_FLOP_FWD      # per direction = per launch
            kname = f"sps_{dom}_kernel<7,1>"
            info = _l.gsp_launch_info(_l.make_gsp_desc(T, B, 0 if kind == "onlysp" else 1))
            # one direction: gx 1024 + gxs 384 + out 512 + stash (1024 gates + 256 states + 512 GRU + 128 qs); bwd: dout + stash reads + adjoints
            by = {"fwd": 4 * (1024 + 384 + 512 + 1024 + 256 + 512 + 128), "bwd": 4 * (512 + 1024 + 2 * 256 + 512 + 128 + 1024 + 2 * 384)}
            by_alg = {"fwd": 800 + 3584, "bwd": 2 * (800 + 3584)}
            din = 1124
            passes = 1
        else:
            flop_utt = SPS_FLOP_BWD if dom == "bwd" else SPS_FLOP_FWD      # per direction = per launch
            kname = f"sps_{dom}_kernel"
            info = _l.sps_launch_info(_l.make_sps_desc(T, B))
            # one direction: gx 1024 + out 512 + stash (2x1024 gates + 5x256 states); bwd: dout 512 + stash reads + 2x1024 adjoints
            by = {"fwd": 4 * (1024 + 512 + 2 * 1024 + 5 * 256 + 4 * 128), "bwd": 4 * (512 + 2 * 1024 + 2 * 256 + 2 * 1024 + 4 * 128)}
            by_alg = {"fwd": 800 + 3584, "bwd": 2 * (800 + 3584)}
            din = 1124
            passes = 1
        flop = flop_utt * T * B
        achieved = flop / (kms[dom] * 1e-3) / 1e12 if kms[dom] > 0 else 0.0
        sm_mhz = (clocks or {}).get("sm_mhz") or pk["sm_max_mhz"]
        ffma_peak = 148 * 128 * 2 * sm_mhz * 1e6 / 1e12
        hbm_gbs = by[dom] * T * B / (kms[dom] * 1e-3) / 1e9 if kms[dom] > 0 else 0.0
        hbm_alg = by_alg[dom] * T * B / (kms[dom] * 1e-3) / 1e9 if kms[dom] > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "dram_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(kname, {}).get("dram_bytes_per_launch")
        kernel_ms = {k: round(v, 4) for k, v in kms.items()}
        shares = {k: v * len(kev[k]) / args.steps / (ms / args.steps) for k, v in kms.items()}
        if kind == "ATV":
            # The dominant kernel runs its products on tcgen05 with the fp32-accurate three-term bf16 split: its roofline is
            # the measured bf16 tensor peak divided by the three passes, in the reference's algorithmic FLOPs.
            tc_peak = pk["bf16_sustained"] / passes
            roof = {"bound": "tensor", "kernel": kname, "achieved": achieved, "peak": tc_peak, "unit": "TFLOP/s",
                    "frac": achieved / tc_peak, "traffic": traffic,
                    "peak_source": f"{pk['source']} bf16 cuBLAS sustained {pk['bf16_sustained']:.0f} TFLOP/s / {passes} split passes",
                    "algorithmic_flop_per_utt": flop_utt,
                    "note": "the step chain is latency-bound (three group exchanges per step through L2), see DESIGN.md §4",
                    "hbm": {"achieved": hbm_alg, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": hbm_alg / pk["hbm_gbs"],
                            "algorithmic_bytes_per_utt": by_alg[dom], "scheduled_bytes_per_utt": by[dom],
                            "scheduled_gbs": hbm_gbs, "scheduled_frac": hbm_gbs / pk["hbm_gbs"]}}
        else:
            roof = {"bound": "hbm", "kernel": kname, "achieved": hbm_gbs, "peak": pk["hbm_gbs"], "unit": "GB/s",
                    "frac": hbm_gbs / pk["hbm_gbs"], "traffic": traffic, "peak_source": pk["source"] + " copy bandwidth",
                    "algorithmic_bytes_per_utt": by[dom],
                    "fp32_ffma": {"achieved": achieved, "peak": ffma_peak, "frac": achieved / ffma_peak,
                                  "note": f"kernel is fp32 FFMA; peak = 148 SM x 128 lanes x 2 x {sm_mhz:.0f} MHz (clock under load)"}}
        roof.update({"kernel_ms": kernel_ms, "launches_per_step": {k: len(v) / args.steps for k, v in kev.items()},
                     "share_of_step": shares,
                     "tensor_core_kernels": {"gemm3": tc_summary(gev), "attention": tc_summary(aev),
                                             "note": "event-timed per launch in 2 extra steps on the single-stream schedule "
                                                     "(outside the timed region, which runs the branches on concurrent streams)"}})
        out = {
            "metric": metric_name(kind), "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": args.dtype, "data": "synthetic",
            "config": workload_config(kind, T, B, world, info),
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": (e2e_ms / args.steps) if e2e_ms else None},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": roof,
        }
        out.update(extras)
        if world == 1 and not args.no_cpu_baseline:
            try:
                out["cpu_baseline"] = run_cpu_baseline(model_kind=kind)
            except Exception as e:
                out["cpu_baseline"] = {"error": repr(e)[:200]}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=BATCH, help="dialogues per GPU")
    ap.add_argument("--seq", type=int, default=T_LEN)
    ap.add_argument("--model", default="ATV", choices=["ATV", "sps", "onlysp", "nsps", "no_en"],
                    help="ATV = BASELINE.json configs[1] (headline); sps = configs[2] shapes (speaker-state model, fp32); "
                         "onlysp / nsps / no_en = the GRU speaker-state variants (train.py's default model, its listener form, "
                         "and the listener form without the text encoder)")
    ap.add_argument("--dtype", default="f32", choices=["f32", "bf16"],
                    help="f32: every tensor-core product is the fp32-accurate split (parity mode, the metric's precision); "
                         "bf16: time-parallel products with bf16 operands (tests/test_bf16_gpu.py states the tolerance)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="profiling runs only: skip the end-to-end leg")
    ap.add_argument("--no-extras", action="store_true",
                    help="skip the extra arms of the N = 1 line (optimizer step, config 3, eager-GPU reference, configs 1 and 4)")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

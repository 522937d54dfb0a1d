"""Development harness for the weight-stationary tensor-core recurrence (lsthm_mab_*): per-tensor errors of the kernel
boundary against the fp64 plain-C oracle, one case per subprocess (a trapped kernel must not take the other cases down),
then kernel timings at the benchmark shape.  Usage (GPU box):  python profiles/dev_mab_check.py [case ...] [--time]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = {
    "atv_t1": ("ATV", 1, 9, 0, False),
    "atv_t2": ("ATV", 2, 5, 0, False),
    "atv_t5": ("ATV", 5, 11, 0, True),
    "at_t3": ("AT", 3, 17, 0, False),
    "at_t5m": ("AT", 5, 11, 0, True),
    "atv_multi": ("ATV", 4, 100, 0, True),        # several groups
    "atv_dg16": ("ATV", 3, 200, 16, False),       # more blocks than groups: waves
    "atv_big": ("ATV", 6, 1024, 0, True),
}


def run_case(name):
    import numpy as np
    import torch
    from importlib import import_module
    import lsthm_b200
    from helpers import SPEC, e_inf, seeded_model
    from oracle import cpu as ocpu
    lib = import_module(lsthm_b200.__name__ + "._lib")
    kind, T, N, rows, masked = CASES[name]
    spec = SPEC[kind]
    dh, rd = spec["dh"], spec["rd"]
    D, MH, M = sum(dh), 64, len(dh)
    seed = T * 100 + N
    model = seeded_model(kind, 100 + seed)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.add_(0.05 * torch.randn(p.shape, generator=g))
    gx = torch.randn(T, N, 4 * D, generator=g)
    mask = (torch.bernoulli(torch.full((T, N, MH), 0.7), generator=g) / 0.7) if masked else None
    dhz = torch.randn(T, N, 2 * D, generator=g)
    weights = [w.detach() for w in model.recurrence_weights()]
    params = {k: v.detach().numpy() for k, v in model.state_dict().items()}
    dev = "cuda"
    w = [t.to(dev).contiguous() for t in weights]
    U, V = w[0:M], w[M:2 * M]
    Watt, batt = w[2 * M], w[2 * M + 1]
    Wr, br = w[2 * M + 2:3 * M + 2], w[3 * M + 2:4 * M + 2]
    Wf1, bf1, Wf2, bf2 = w[4 * M + 2:4 * M + 6]
    d = lib.make_desc(T, N, dh, rd, MH, 4, rows)
    ws = lib.make_weights(U, V, Watt, batt, Wr, br, Wf1, bf1, Wf2, bf2)
    print(name, "launch:", lib.mab_launch_info(d), flush=True)
    packed = torch.empty(lib.mab_pack_bytes(d), device=dev, dtype=torch.uint8)
    lib.mab_pack(d, ws, packed)
    work = torch.empty(lib.mab_workspace_bytes(d), device=dev, dtype=torch.uint8)
    new = lambda *s: torch.full(s, float("nan"), device=dev)
    hz, UH = new(T, N, 2 * D), new(T, N, MH)
    sC = new(T, N, D)
    st = lib.mab_alloc_stash(d, dev)
    for v in st.values():
        v.fill_(float("nan"))
    gxc = gx.to(dev)
    mc = None if mask is None else mask.to(dev)
    lib.mab_fwd(d, packed, gxc, mc, hz, UH, sC, st["sCp"], st["sG"], st["sE"], st["sMS"], st["sP"], work)
    torch.cuda.synchronize()
    sG = lib.mab_unblock(st["sG"], d, 4 * D)
    sE = lib.mab_unblock(st["sE"], d, 4 * D)
    sP = lib.mab_unblock(st["sP"], d, 4 * MH).view(T, N, 4, MH)
    sMS = lib.mab_unblock(st["sMS"], d, 8).view(T, N, 4, 2)
    assert torch.equal(lib.mab_unblock(st["sCp"], d, D), sC), "private c copy differs"
    mask64 = None if mask is None else mask.double().numpy()
    ref = ocpu.mab_forward(params, gx.double().numpy(), dh, rd, mask64)
    A = (torch.exp(sE.view(T, N, 4, D) - sMS[..., 0:1]) * sMS[..., 1:2])
    out = dict(H=hz[:, :, :D], C=sC, G=sG, A=A, UH=UH)
    refs = dict(H=ref["hz"][:, :, :D], C=ref["C"], G=ref["G"], A=ref["A"], UH=ref["UH"])
    # P_k = W1[:, head block] . (a_k * c) from the oracle's A and C, composite W1 in fp64
    Wf1_64 = np.asarray(params["fc.0.weight"], np.float64)
    W1 = np.zeros((MH, 4, D))
    o = ro = 0
    for mi, h in enumerate(dh):
        Wr64 = np.asarray(params[f"reduce_dim_nn_{'lav'[mi]}.0.weight"], np.float64).reshape(rd[mi], 4, h)
        W1[:, :, o:o + h] = np.einsum("qr,rkj->qkj", Wf1_64[:, ro:ro + rd[mi]], Wr64)
        o += h
        ro += rd[mi]
    refs["P"] = np.einsum("qkj,tnkj->tnkq", W1, ref["A"] * ref["C"][:, :, None, :])
    out["P"] = sP
    ferr = {k: e_inf(out[k].cpu().numpy().reshape(refs[k].shape), refs[k]) for k in out}
    print(name, "FWD", {k: f"{v:.2e}" for k, v in ferr.items()}, flush=True)
    # backward on the kernel's own stash
    dhzc = dhz.to(dev)
    duz = (dhzc[:, :, D:] @ Wf2).contiguous()
    dgx, de, dup, att = new(T, N, 4 * D), new(T, N, 4 * D), new(T, N, MH), new(T, N, 4 * D)
    lib.mab_bwd(d, packed, dhzc, duz, mc, st["sCp"], st["sG"], st["sE"], st["sMS"], st["sP"], UH, dgx, de, dup, att, work)
    torch.cuda.synchronize()
    radj, _ = ocpu.mab_backward(params, dhz.double().numpy(), ref, dh, rd, mask64)
    adj = dict(dgx=dgx, de=de, dup=dup)
    berr = {k: e_inf(adj[k].cpu().numpy().reshape(radj[k].shape), radj[k]) for k in adj}
    a4 = A * sC.view(T, N, 1, D)
    o = 0
    att_ok = True
    for mi, h in enumerate(dh):
        att_ok &= bool(torch.allclose(att[:, :, 4 * o:4 * o + 4 * h], a4[:, :, :, o:o + h].reshape(T, N, 4 * h), rtol=1e-5, atol=1e-7))
        o += h
    print(name, "BWD", {k: f"{v:.2e}" for k, v in berr.items()}, "att_ok", att_ok, flush=True)
    ok = all(np.isfinite(v) and v < 2e-5 for v in ferr.values()) and all(np.isfinite(v) and v < 1e-4 for v in berr.values()) and att_ok
    print(name, "OK" if ok else "FAIL", flush=True)
    return 0 if ok else 1


def run_timing():
    import torch
    from importlib import import_module
    import lsthm_b200
    from helpers import SPEC, seeded_model
    lib = import_module(lsthm_b200.__name__ + "._lib")
    prof = os.environ.get("MAB2_PROFILE") == "1"          # under ncu: one configuration, two launches of each kernel
    for kind, T, N in ((("ATV", 110, 1024),) if prof else (("ATV", 110, 1024), ("ATV", 110, 128), ("AT", 110, 1024))):
        spec = SPEC[kind]
        dh, rd = spec["dh"], spec["rd"]
        D, MH, M = sum(dh), 64, len(dh)
        model = seeded_model(kind, 111)
        dev = "cuda"
        w = [t.detach().to(dev).contiguous() for t in model.recurrence_weights()]
        U, V = w[0:M], w[M:2 * M]
        Watt, batt = w[2 * M], w[2 * M + 1]
        Wr, br = w[2 * M + 2:3 * M + 2], w[3 * M + 2:4 * M + 2]
        Wf1, bf1, Wf2, bf2 = w[4 * M + 2:4 * M + 6]
        d = lib.make_desc(T, N, dh, rd, MH, 4, 0)
        ws = lib.make_weights(U, V, Watt, batt, Wr, br, Wf1, bf1, Wf2, bf2)
        packed = torch.empty(lib.mab_pack_bytes(d), device=dev, dtype=torch.uint8)
        work = torch.empty(lib.mab_workspace_bytes(d), device=dev, dtype=torch.uint8)
        new = lambda *s: torch.empty(s, device=dev)
        gx = torch.randn(T, N, 4 * D, device=dev)
        mask = torch.bernoulli(torch.full((T, N, MH), 0.7, device=dev)) / 0.7
        hz, UH = new(T, N, 2 * D), new(T, N, MH)
        sC = new(T, N, D)
        st = lib.mab_alloc_stash(d, dev)
        sCp, sG, sE, sMS, sP = st["sCp"], st["sG"], st["sE"], st["sMS"], st["sP"]
        dhz = torch.randn(T, N, 2 * D, device=dev)
        duz = torch.randn(T, N, MH, device=dev)
        dgx, de, dup, att = new(T, N, 4 * D), new(T, N, 4 * D), new(T, N, MH), new(T, N, 4 * D)
        res = {}
        for what in ("pack", "fwd", "bwd"):
            fn = {"pack": lambda: lib.mab_pack(d, ws, packed),
                  "fwd": lambda: lib.mab_fwd(d, packed, gx, mask, hz, UH, sC, sCp, sG, sE, sMS, sP, work),
                  "bwd": lambda: lib.mab_bwd(d, packed, dhz, duz, mask, sCp, sG, sE, sMS, sP, UH, dgx, de, dup, att, work)}[what]
            for _ in range(1 if prof else 3):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(1 if prof else 10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[what] = e0.elapsed_time(e1) / 10
        print("TIMING", kind, T, N, lib.mab_launch_info(d), {k: f"{v:.3f} ms" for k, v in res.items()}, flush=True)
        if prof:
            continue
        # phase trace of block 0 (control thread / first epilogue warp), cycles, median over the steps
        tr = torch.zeros(T, 2, 16, device=dev, dtype=torch.int64)
        for what in ("fwd", "bwd"):
            tr.zero_()
            lib.mab_set_trace(tr)
            if what == "fwd":
                lib.mab_fwd(d, packed, gx, mask, hz, UH, sC, sCp, sG, sE, sMS, sP, work)
            else:
                lib.mab_bwd(d, packed, dhz, duz, mask, sCp, sG, sE, sMS, sP, UH, dgx, de, dup, att, work)
            torch.cuda.synchronize()
            lib.mab_set_trace(None)
            tc = tr.cpu()
            if int(tc.abs().sum()) == 0:
                continue
            for role in (0, 1):
                x = tc[5:T - 5, role, :]                      # steady-state steps
                nz = [k for k in range(16) if int(x[:, k].abs().sum()) > 0]
                if not nz:
                    continue
                base = x[:, nz[0]]
                rel = {k: int((x[:, k] - base).median()) for k in nz}
                step = int((tc[6:T - 4, role, nz[0]] - tc[5:T - 5, role, nz[0]]).median())
                print(f"TRACE {what} {kind} N={N} role{role} step={step} cyc:", rel, flush=True)


if __name__ == "__main__":
    args = sys.argv[1:]
    if args and args[0] == "--case":
        sys.exit(run_case(args[1]))
    if args and args[0] == "--timing":
        run_timing()
        sys.exit(0)
    names = [a for a in args if not a.startswith("--")] or list(CASES)
    bad = 0
    for n in names:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, __file__, "--case", n], timeout=300)
            rc = r.returncode
        except subprocess.TimeoutExpired:
            rc = -999
        print(f"[case {n}] rc={rc} {time.time() - t0:.1f}s", flush=True)
        bad += rc != 0
        if rc not in (0, 1):
            print("stopping: a case crashed or timed out", flush=True)
            break
    if "--time" in args and bad == 0:
        subprocess.run([sys.executable, __file__, "--timing"], timeout=600)
    sys.exit(1 if bad else 0)

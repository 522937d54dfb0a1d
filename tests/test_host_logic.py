"""CPU suite, part 2: the host side above the C ABI, with the CUDA library replaced by the plain-C
oracle (tests/oracle_backend.py).  Checks the drop-in modules end to end against the
reference-generated fixtures: module wiring, autograd.Function plumbing and the hoisted
weight-gradient products of recurrence.py.  Also: the product path refuses to run without CUDA."""
import pytest
from importlib import import_module
import lsthm_b200
import torch

import oracle_backend
from helpers import SPEC, check_against_golden, golden_files, load_golden, run_module, seeded_model


@pytest.mark.parametrize("path", golden_files(), ids=lambda p: p.split("/")[-1][:-4])
def test_module_with_oracle_backend_matches_fixture(path, monkeypatch):
    oracle_backend.install(monkeypatch)
    fix = load_golden(path)
    probs, loss, dx, grads = run_module(fix)
    check_against_golden(fix, probs, loss, dx, grads, tol_out=5e-6, tol_grad=5e-5)


def test_product_path_refuses_cpu():
    model = seeded_model("AT", 1).eval()
    with pytest.raises(RuntimeError, match="CUDA"):
        model(torch.randn(3, 2, SPEC["AT"]["din"]))


def test_unused_parameters_stay_gradless(monkeypatch):
    oracle_backend.install(monkeypatch)
    model = seeded_model("AT", 1).eval()
    model(torch.randn(3, 2, 200)).sum().backward()
    none = [n for n, p in model.named_parameters() if p.grad is None]
    assert sorted(none) == sorted(f"encoder_{m}.pos_ffn.fc.{w}" for m in "la" for w in ("weight", "bias"))


def test_every_kernel_binding_refuses_host_tensors():
    """No CPU fallback anywhere on the product path: each binding raises on host tensors before touching the library."""
    from importlib import import_module

    import lsthm_b200
    lib = import_module(lsthm_b200.__name__ + "._lib")
    fd = import_module(lsthm_b200.__name__ + ".fused_dln")
    pl = import_module(lsthm_b200.__name__ + ".pipeline")
    fa = import_module(lsthm_b200.__name__ + ".fused_attention")
    a, b = torch.randn(8, 8), torch.randn(8, 8)
    with pytest.raises(RuntimeError):
        lib.gemm3(lib.GEMM_NT, a, b)
    with pytest.raises(RuntimeError):
        lib.colsum(a)
    with pytest.raises(RuntimeError):
        fd.drop_res_layer_norm(a, None, b, torch.ones(8), torch.zeros(8), 1e-6)
    with pytest.raises(RuntimeError):
        pl.assemble_input(*(torch.zeros(2, 2, 8) for _ in range(5)))
    with pytest.raises(RuntimeError):
        fa.fused_self_attention(torch.randn(1, 4, 960), 8, 40 ** -0.5)
    with pytest.raises(RuntimeError):
        lib.adam_step(a.view(-1), b.view(-1), a.view(-1).clone(), a.view(-1).clone(), 1e-3, 0.9, 0.999, 1e-8, 0.0, 1)


def test_precision_switch_validates_and_sets_the_descriptor():
    from importlib import import_module

    import lsthm_b200
    lib = import_module(lsthm_b200.__name__ + "._lib")
    with pytest.raises(ValueError):
        lib.set_precision("fp16")
    try:
        lib.set_precision("bf16")
        assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1).precision == 1
        assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1, time_major=True).row_stride_i == 2
    finally:
        lib.set_precision("fp32")
    assert lib.make_attn_desc(2, 5, 8, 960, 960, 960, 320, 0.1).precision == 0


def test_rows_view_finds_copy_free_row_matrices():
    from importlib import import_module

    import lsthm_b200
    mm3 = import_module(lsthm_b200.__name__ + ".mm3")
    x = torch.randn(7, 3, 24)                                  # time-major storage [L, B, K]
    rows, swapped = mm3.rows_view(x)
    assert not swapped and rows.data_ptr() == x.data_ptr() and rows.shape == (21, 24)
    sl = x[:, :, 4:12].permute(1, 0, 2)                        # the view the reference hands to its encoders
    rows, swapped = mm3.rows_view(sl)
    assert swapped and rows.shape == (21, 8) and rows.stride() == (24, 1)
    assert torch.equal(rows.view(7, 3, 8), x[:, :, 4:12])
    assert mm3.rows_view(x.permute(2, 0, 1)) is None           # no unit inner stride -> caller copies


# ------------------------------------------------------------------------------------------------
# data-parallel hookup: never-used parameters, bucket views, optimizer state
# ------------------------------------------------------------------------------------------------
def _ddp():
    from importlib import import_module
    return import_module(lsthm_b200.__name__ + ".ddp")


@pytest.mark.parametrize("pattern", ["mab_*.npz", "sps_*.npz", "onlysp_*.npz", "nsps_*.npz", "no_en_*.npz"])
def test_unused_parameter_list_equals_the_reference_grad_none_set(pattern):
    """the reducer must leave exactly the parameters whose grad is None in the REFERENCE step out of its
    buckets (20 tensors for MARN1_sps, 38 / 22 / 34 for onlysp / nsps / no_en).  The fixtures record that set (gnone/*) from
    the live reference."""
    from helpers import SPEAKER_MODELS, golden_files, load_golden
    files = golden_files(pattern)
    assert files
    for f in files:
        fix = load_golden(f)
        kind = str(fix["kind"]) if "kind" in fix else "sps"
        model = {"ATV": lambda: lsthm_b200.HybridRNN_ATV.MARN(), "AT": lambda: lsthm_b200.HybridRNN_AT.MARN()}.get(
            kind, SPEAKER_MODELS.get(kind))()
        assert sorted(_ddp().unused_parameter_names(model)) == sorted(k[6:] for k in fix if k.startswith("gnone/")), f
        if kind not in ("ATV", "AT"):
            assert len(_ddp().unused_parameter_names(model)) == {"sps": 20, "onlysp": 38, "nsps": 22, "no_en": 34}[kind]
            red = _ddp().GradAllReducer(model, 1)
            bucketed = {id(p) for members in red._members for p in members}
            for n, p in model.named_parameters():
                assert (id(p) not in bucketed) == (n in red.skipped), n      # never-used parameters have no bucket slot


def test_reducer_packs_whatever_autograd_produced_and_rejects_a_missing_rearm():
    """Gradients start as None (autograd hands its tensors over, no accumulate-add launches); when a bucket is complete they
    are packed into the flat buffer by one multi-tensor copy and ``p.grad`` become views of it.  A gradient that already IS
    the view (someone zeroed in place instead of dropping) is left where it is; a backward without ``reducer.zero_grad()``
    in between is an error, never a silent reduction of stale buckets."""
    torch.manual_seed(0)
    model = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    red = _ddp().GradAllReducer(model, 1, bucket_bytes=64)
    assert all(p.grad is None for p in model.parameters())
    x = torch.randn(3, 8)
    model(x).sum().backward()
    red.finish()
    g0 = [p.grad.clone() for p in model.parameters()]
    assert all(p.grad.data_ptr() == red._expected_ptr(p) for p in model.parameters())
    ref = torch.nn.Sequential(torch.nn.Linear(8, 8), torch.nn.Linear(8, 4))
    ref.load_state_dict(model.state_dict())
    ref(x).sum().backward()
    assert all(torch.equal(a.grad, b.grad) for a, b in zip(model.parameters(), ref.parameters()))
    assert sorted(red.fire_order) == list(range(len(red.buckets))) and len(red.buckets) >= 2
    with pytest.raises(RuntimeError, match="more gradients than it has members"):
        model(x).sum().backward()                                                # second backward without zero_grad
    red.zero_grad()
    for p in model.parameters():                                                 # in-place zeroed views instead of None
        p.grad = red._view[id(p)].zero_()
    model(x).sum().backward()
    red.finish()
    for p, g in zip(model.parameters(), g0):
        assert p.grad.data_ptr() == red._expected_ptr(p) and torch.allclose(p.grad, g)
    # a parameter without a gradient in this step: zero slot, grad stays None, the others are unaffected
    red.zero_grad()
    model[1](x[:, :8]).sum().backward()
    red.finish()
    assert model[0].weight.grad is None and float(red._view[id(model[0].weight)].abs().sum()) == 0.0
    assert torch.allclose(model[1].bias.grad, g0[3])


def test_fused_adam_exposes_param_groups_and_state(monkeypatch):
    """StepLR (model_trainer.py:83) attaches to FusedAdam and its lr reaches the kernel call; state round-trips."""
    torch.manual_seed(0)
    model = torch.nn.Linear(6, 5)
    ddp = _ddp()
    red = ddp.GradAllReducer(model, 1, flatten_params=True)
    calls = []
    lib = import_module(lsthm_b200.__name__ + "._lib")
    monkeypatch.setattr(lib, "adam_step", lambda p, g, m, v, lr, b1, b2, eps, wd, step: calls.append((lr, wd, step)))
    opt = ddp.FusedAdam(red, lr=1e-3, weight_decay=2e-5)
    sched = torch.optim.lr_scheduler.StepLR(opt, step_size=1, gamma=0.98)
    opt.step(); sched.step(); opt.step()
    assert calls[0][0] == pytest.approx(1e-3) and calls[-1][0] == pytest.approx(0.98e-3) and calls[-1][1] == 2e-5
    assert calls[-1][2] == 2
    sd = opt.state_dict()
    opt2 = ddp.FusedAdam(red, lr=5e-2)
    opt2.load_state_dict(sd)
    assert opt2.step_count == 2 and opt2.lr == pytest.approx(0.98e-3) and opt2.param_groups[0]["weight_decay"] == 2e-5


def test_length_bucketed_batching_and_pinned_collate():
    """bucketing covers every dialogue once, cuts padding on the res.csv-like length distribution, keeps
    the shards of one step at one padded length (F11); the collate equals the reference's pad_sequence collate_fn."""
    from importlib import import_module
    from torch.nn.utils.rnn import pad_sequence
    pl = import_module(lsthm_b200.__name__ + ".pipeline")
    g = torch.Generator().manual_seed(0)
    lens = torch.clamp(torch.round(52.4 + 17.4 * torch.randn(600, generator=g)), 8, 110).int().tolist()
    bs = pl.LengthBucketBatchSampler(lens, 32, pool_batches=8, seed=1)
    seen = sorted(i for b in bs for i in b)
    assert seen == list(range(600)) and len(bs) == 19
    rnd = pl.LengthBucketBatchSampler(lens, 32, pool_batches=1, seed=1)      # pool of one batch == plain random batches
    assert bs.padding_fraction() < 0.5 * rnd.padding_fraction() and rnd.padding_fraction() > 0.25
    bs.set_epoch(1)
    assert sorted(i for b in bs for i in b) == list(range(600)) and [b for b in bs] != [b for b in rnd]
    # two ranks: disjoint shards of every global batch, same padded length per step
    r0 = pl.LengthBucketBatchSampler(lens, 32, seed=3, world=2, rank=0)
    r1 = pl.LengthBucketBatchSampler(lens, 32, seed=3, world=2, rank=1)
    for step, (a, b) in enumerate(zip(r0, r1)):
        assert not set(a) & set(b) and abs(len(a) - len(b)) <= 1
        assert r0.global_max_len(step) == r1.global_max_len(step) >= max(lens[i] for i in a + b)
    assert sorted(i for s in (r0, r1) for b in s for i in b) == list(range(600))
    # collate == dataloader.py:45-47 (pad_sequence time-major for fields < 7, batch-major for umask/label, ids as a list)
    def sample(n, vid):
        return (torch.randn(n, 8), torch.randn(n, 8), torch.randn(n, 8), torch.randn(n, 8), torch.randn(n, 5), torch.randn(n, 3),
                torch.eye(2)[torch.randint(0, 2, (n,))], torch.ones(n), torch.randint(0, 6, (n,)), vid)
    data = [sample(n, f"d{n}") for n in (5, 9, 2)]
    got = pl.collate_dialogues(data)
    for f in range(10):
        col = [s[f] for s in data]
        want = pad_sequence(col) if f < 7 else pad_sequence(col, True) if f < 9 else col
        assert (got[f] == want) if f == 9 else torch.equal(got[f], want), f
    assert pl.collate_dialogues(data, pad_to=12)[0].shape == (12, 3, 8) and pl.collate_dialogues(data, pad_to=12)[7].shape == (3, 12)


def test_single_step_cell_signatures_match_the_oracle():
    """API parity of the container cells: LSTHM / LSTHM1 / in-cell CrossAttention keep the reference's single-step
    ``forward`` signatures; values equal the oracle's restatement of the same lines."""
    from oracle import torch_port as tp
    torch.manual_seed(3)
    net = lsthm_b200.HybridRNN_ATV.MARN()
    p = {k: v.detach() for k, v in net.state_dict().items()}
    x, c, h, z = torch.randn(5, 100), torch.randn(5, 128), torch.randn(5, 128), torch.randn(5, 208)
    c1, h1 = net.lsthm_l(x, c, h, z)
    c2, h2 = tp.lsthm_cell(p, "lsthm_l", x, c, h, z)
    assert torch.allclose(c1, c2, atol=1e-6) and torch.allclose(h1, h2, atol=1e-6)
    sps = lsthm_b200.lsthm_sps.MARN1_sps(6).eval()
    tp.perturb_ones(sps, 4)
    ps = {k: v.detach() for k, v in sps.state_dict().items()}
    cell = sps.marn_cell_f
    zz, s = torch.randn(5, 128), torch.randn(5, 128)
    c1, h1 = cell.lsthm_l(x, c, h, zz, s)
    c2, h2 = tp.lsthm1_cell(ps, "marn_cell_f.lsthm_l", x, c, h, zz, s)
    assert torch.allclose(c1, c2, atol=1e-6) and torch.allclose(h1, h2, atol=1e-6)
    a1 = cell.crossatt_l2a(c, h)
    a2 = tp.cross_attention_cell(ps, "marn_cell_f.crossatt_l2a", c, h, None, "")
    assert torch.allclose(a1, a2, atol=2e-6)


def test_modules_copy_and_pickle_without_their_cached_streams():
    """The drop-in modules cache CUDA streams on themselves (mab_net / streams.fork_join); ``copy.deepcopy`` and ``pickle`` of a
    module (EMA copies, ``torch.save(model)``) must not try to copy them.  A lock stands in for a stream here (neither pickles)."""
    import copy
    import pickle
    import threading
    m = lsthm_b200.HybridRNN_AT.MARN()
    m._side_streams[0] = [threading.Lock()]
    s = lsthm_b200.lsthm_onlysp.MARN1_onlysp(6)
    s.__dict__["_fork_streams"] = {0: [threading.Lock()]}
    for mod, attr in ((m, "_side_streams"), (s, "_fork_streams")):
        for clone in (copy.deepcopy(mod), pickle.loads(pickle.dumps(mod))):
            assert getattr(clone, attr) == {} and len(getattr(mod, attr)) == 1
            a, b = mod.state_dict(), clone.state_dict()
            assert list(a) == list(b) and all(torch.equal(a[k], b[k]) for k in a)
            assert clone.concurrent_encoders is True

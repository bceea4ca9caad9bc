"""Probe heads on the GPU (-m gpu): fused forward + argmax vs the oracle, the native train step vs the reference's
torch AdamW step, and shard-sum equivalence of the data-parallel phases (ranks emulated on one GPU)."""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import probe_oracle, tower_oracle as O  # noqa: E402
from vlmfw_b200 import probes  # noqa: E402
from vlmfw_b200.probe_train import ProbeTrainStep, device_keep_mask, mask_key  # noqa: E402

LOGIT_TOL = 1e-2  # BASELINE.json north_star: probe-logit max abs error


def _head(dim, classes, hidden, seed):
    p = O.seeded_head_params(dim, classes, hidden, seed=seed)
    m = (probes.make_head_deeper(dim, classes, hidden, 0.3) if hidden else probes.make_head(dim, classes, 0.3))
    m.load_state_dict(p)
    return m.cuda().eval(), p


@pytest.mark.parametrize("dim", [1152, 1024, 1408, 144])
@pytest.mark.parametrize("batch", [1, 7, 256])
def test_fused_heads_forward(cuda_device, dim, batch):
    feats = torch.randn(batch, dim, generator=torch.Generator().manual_seed(dim + batch))
    cfgs = [(2, 0), (12, 0), (9, 0), (7, 512), (2, 64)]
    mods, params = zip(*[_head(dim, c, h, 300 + i) for i, (c, h) in enumerate(cfgs)])
    logits, amax = probes.fused_heads_forward(feats.cuda(), list(mods))
    for i, p in enumerate(params):
        ref = O.head_forward(feats, p)
        torch.testing.assert_close(logits[i].cpu(), ref, rtol=1e-4, atol=1e-4)  # fp32 both sides
        assert amax[:, i].cpu().tolist() == ref.argmax(1).tolist()  # bit-exact labels


def test_argmax_first_max_tie_rule(cuda_device):
    m = probes.make_head(8, 4, 0.0)
    with torch.no_grad():
        m[2].weight.zero_()
        m[2].bias.copy_(torch.tensor([1.0, 3.0, 3.0, 2.0]))
    m = m.cuda().eval()
    _, amax = probes.fused_heads_forward(torch.randn(5, 8).cuda(), [m])
    assert amax[:, 0].tolist() == [1] * 5 == torch.tensor([1.0, 3.0, 3.0, 2.0]).argmax().repeat(5).tolist()


@pytest.mark.parametrize("dim,classes,n,p", [(1152, 9, 256, 0.3), (144, 5, 33, 0.0), (1024, 2, 64, 0.5), (256, 9, 2048, 0.3)])
def test_train_steps_match_reference_adamw(cuda_device, dim, classes, n, p):
    g = torch.Generator().manual_seed(n)
    state = O.seeded_head_params(dim, classes, 0, seed=9)
    cw = torch.rand(classes, generator=g) + 0.5
    batches = []
    for _ in range(3):
        x = torch.randn(n, dim, generator=g) * 1.5 + 0.2
        y = torch.randint(0, classes, (n,), generator=g)
        y[::9] = -1  # ignore_index rows
        keep = (torch.rand(n, dim, generator=g) >= p).to(torch.uint8) if p > 0 else None
        batches.append((x, y, keep))
    ref_state, ref_losses = probe_oracle.reference_train_steps(state, batches, classes, p, cw, lr=1e-3, weight_decay=1e-2)
    ts = ProbeTrainStep(dim, classes, "cuda", lr=1e-3, weight_decay=1e-2, dropout_p=p, class_weight=cw)
    ts.load_head_state_dict({k: v.cuda() for k, v in state.items()})
    losses = [ts.step(x.cuda(), y.cuda(), keep_mask=None if k is None else k.cuda()) for x, y, k in batches]
    for a, b in zip(losses, ref_losses):
        assert abs(a - b) < 2e-5 * max(1.0, abs(b))
    got = ts.head_state_dict()
    for k in ("2.weight", "2.bias", "0.weight", "0.bias", "0.running_mean", "0.running_var"):
        torch.testing.assert_close(got[k].cpu(), ref_state[k], rtol=2e-4, atol=2e-6, msg=lambda m, k=k: f"{k}: {m}")
    assert int(got["0.num_batches_tracked"]) == int(state["0.num_batches_tracked"]) + 3


def test_sharded_phases_equal_single_process_gradient(cuda_device):
    """Config 5's equality check: all-reduced gradient == single-process gradient on the concatenated batch."""
    dim, classes, n, p = 1152, 9, 256, 0.3
    g = torch.Generator().manual_seed(1)
    state = {k: v.cuda() for k, v in O.seeded_head_params(dim, classes, 0, seed=4).items()}
    x = (torch.randn(n, dim, generator=g) * 1.5).cuda()
    y = torch.randint(0, classes, (n,), generator=g)
    y[::11] = -1
    y = y.cuda()
    single = ProbeTrainStep(dim, classes, "cuda", dropout_p=p, seed=5)
    single.load_head_state_dict(state)
    keep = single.dropout_keep_mask(n, 0, n)
    single.phase1_stats(x)
    full_bucket = single.phase2_grads(x, y, n, keep).clone()
    # emulate 4 ranks: each computes its phases on its slice; sums stand in for the NCCL all-reduce
    from vlmfw_b200.parallel import shard_range

    ranks = []
    for r in range(4):
        t = ProbeTrainStep(dim, classes, "cuda", dropout_p=p, seed=5)
        t.load_head_state_dict(state)
        ranks.append(t)
    spans = [shard_range(n, r, 4) for r in range(4)]
    stats = sum(t.phase1_stats(x[lo:hi]).clone() for t, (lo, hi) in zip(ranks, spans))
    torch.testing.assert_close(stats, single.stats, rtol=1e-12, atol=1e-9)
    total = torch.zeros_like(full_bucket)
    for t, (lo, hi) in zip(ranks, spans):
        t.stats.copy_(stats)
        assert torch.equal(t.dropout_keep_mask(n, lo, hi), keep[lo:hi])  # mask keyed by global row
        total += t.phase2_grads(x[lo:hi], y[lo:hi], n, keep[lo:hi])
    torch.testing.assert_close(total, full_bucket, rtol=1e-4, atol=1e-5)
    # and both agree with the fp64 oracle gradient
    ref = probe_oracle.phase2_bucket(x.cpu(), y.cpu(), single.stats.cpu(), n, state["0.weight"].cpu(),
                                     state["0.bias"].cpu(), state["2.weight"].cpu(), state["2.bias"].cpu(), None,
                                     keep.cpu(), p)
    torch.testing.assert_close(total.cpu().double(), ref, rtol=1e-3, atol=1e-4)


@pytest.mark.parametrize("rows,cols,row0,p", [(256, 1152, 0, 0.3), (7, 144, 1000003, 0.5), (1, 5, 0, 0.3),
                                               (33, 512, 12345678901, 0.1), (0, 1152, 4, 0.3), (64, 1408, 192, 0.0),
                                               (3, 17, 2, 0.999)])
def test_device_keep_mask_matches_generator_restatement(cuda_device, rows, cols, row0, p):
    """The device dropout mask equals its numpy restatement bit for bit (ragged tails, 64-bit element indices)."""
    from vlmfw_b200 import _lib

    key = mask_key(7, 41, 3)
    got = device_keep_mask(_lib.load(), cuda_device, rows, cols, row0, key, p).cpu()
    assert torch.equal(got, probe_oracle.keep_mask_reference(rows, cols, row0, key, p))
    if p == 0.0:
        assert bool(got.all())


def test_device_keep_mask_is_shard_invariant_and_bernoulli(cuda_device):
    from vlmfw_b200 import _lib

    lib, n, d, p = _lib.load(), 2048, 1152, 0.3
    full = device_keep_mask(lib, cuda_device, n, d, 0, mask_key(1, 5), p)
    for lo, hi in [(0, 683), (683, 1366), (1366, 2048), (5, 6), (2047, 2048)]:
        assert torch.equal(device_keep_mask(lib, cuda_device, hi - lo, d, lo, mask_key(1, 5), p), full[lo:hi])
    rate = full.float().mean().item()
    assert abs(rate - (1 - p)) < 4 * (p * (1 - p) / (n * d)) ** 0.5 + 1e-4  # 4 sigma of a Bernoulli mean
    # per-column and per-row keep rates show no structure
    assert (full.float().mean(0) - (1 - p)).abs().max().item() < 0.06
    assert (full.float().mean(1) - (1 - p)).abs().max().item() < 0.08
    other = device_keep_mask(lib, cuda_device, n, d, 0, mask_key(1, 6), p)  # next step: a fresh draw
    agree = (other == full).float().mean().item()
    assert abs(agree - (p * p + (1 - p) * (1 - p))) < 0.01
    ts = ProbeTrainStep(d, 9, device=cuda_device, dropout_p=p, seed=1)
    ts.step_count = 5
    assert torch.equal(ts.dropout_keep_mask(n, 100, 356), full[100:356])


@pytest.mark.parametrize("comm", ["auto", "p2p"])
@pytest.mark.parametrize("p", [0.3, 0.0])
def test_graph_step_equals_eager_step(cuda_device, p, comm):
    """The train step replayed from one CUDA graph (device-side step counter for AdamW's bias corrections and the dropout
    key) = the eager step: same masks bit for bit, same losses and parameters (bias corrections come from the device's
    powf instead of the host's), also when eager and graph steps alternate. comm="p2p" runs the fused-exchange form
    (vlmfw_probe_linear_step_dp) as a one-rank communicator: its own slots, flags and epoch counter, same arithmetic."""
    d, c, n, steps = 1152, 9, 256, 6
    g = torch.Generator().manual_seed(5)
    cw = torch.rand(c, generator=g) + 0.5
    xs = [(torch.randn(n, d, generator=g) * 1.3 + 0.2).cuda() for _ in range(steps)]
    ys = [torch.randint(-1, c, (n,), generator=g).cuda() for _ in range(steps)]
    eager = ProbeTrainStep(d, c, device=cuda_device, class_weight=cw, dropout_p=p, seed=3)
    graph = ProbeTrainStep(d, c, device=cuda_device, class_weight=cw, dropout_p=p, seed=3, comm=comm)
    mixed = ProbeTrainStep(d, c, device=cuda_device, class_weight=cw, dropout_p=p, seed=3, comm=comm)
    graph.capture(n)
    with pytest.raises(ValueError):
        graph.step_graph(xs[0][:7], ys[0][:7])
    mixed.step(xs[0], ys[0])  # one eager step BEFORE capture: the counter starts at 1
    mixed.capture(n)
    for s in range(steps):
        le = eager.step(xs[s], ys[s])
        if p > 0:
            # the mask the graph is about to draw from the device counter = the eager mask of this step
            assert graph.step_count == s
        lg = graph.step_graph(xs[s].cpu().pin_memory() if s % 2 else xs[s], ys[s])
        if p > 0:
            eager.step_count -= 1
            assert torch.equal(graph._g_keep, eager.dropout_keep_mask(n, 0, n))
            eager.step_count += 1
        assert abs(le - lg) <= 1e-6 * max(1.0, abs(le))
        if s >= 1:
            lm = mixed.step_graph(xs[s], ys[s]) if s % 2 else mixed.step(xs[s], ys[s])
            assert abs(le - lm) <= 1e-6 * max(1.0, abs(le))
    for other in (graph, mixed):
        assert other.step_count == eager.step_count == steps
        torch.testing.assert_close(other.params, eager.params, rtol=0, atol=2e-7)
        torch.testing.assert_close(other.exp_avg_sq, eager.exp_avg_sq, rtol=1e-6, atol=1e-12)
        torch.testing.assert_close(other.running_var, eager.running_var, rtol=1e-6, atol=1e-7)
    torch.testing.assert_close(graph.last_logits, eager.last_logits, rtol=1e-5, atol=1e-6)
    for other in (graph, mixed):
        assert not other.comm_error()
        other.release_graph()


def test_all_ignored_batch_skips_the_update(cuda_device):
    """Every target -1: the reference's loss is 0/0 (NaN into the head); this step reports loss 0.0 and leaves parameters
    and moments untouched, eager and graph form alike (documented deviation, DESIGN.md numerics)."""
    d, c, n = 64, 5, 16
    x = torch.randn(n, d, generator=torch.Generator().manual_seed(0)).cuda()
    y = torch.full((n,), -1, dtype=torch.int64).cuda()
    for comm in ("auto", "p2p"):
        ts = ProbeTrainStep(d, c, device=cuda_device, dropout_p=0.0, seed=3, comm=comm)
        before = ts.params.clone()
        assert ts.step(x, y) == 0.0
        ts.capture(n)
        assert ts.step_graph(x, y) == 0.0
        assert torch.equal(ts.params, before) and float(ts.exp_avg_sq.abs().max()) == 0.0
        ts.release_graph()

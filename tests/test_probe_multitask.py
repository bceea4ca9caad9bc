"""Deeper-head / multitask probe train step: the multitask oracle against the single-task one (CPU), the
CUDA step against the oracle (torch modules + autograd + AdamW in fp64), shard-sum equivalence, unlabeled heads and the
EMA task weights of the reference's RunningMeans."""
import pytest
import torch

from oracle import probe_oracle, tower_oracle as O  # noqa: E402

MIVIA = {"gender": 2, "upper": 12, "lower": 12, "bag": 2, "hat": 2}


def _batches(dim, tasks, hidden, n, p, steps, seed, missing=0.2):
    g = torch.Generator().manual_seed(seed)
    out = []
    for _ in range(steps):
        x = torch.randn(n, dim, generator=g) * 1.5 + 0.2
        y = torch.stack([torch.randint(0, c, (n,), generator=g) for c in tasks.values()], dim=1)
        y[torch.rand(n, len(tasks), generator=g) < missing] = -1
        masks = []
        for _t in tasks:
            k1 = (torch.rand(n, dim, generator=g) >= p).to(torch.uint8) if p > 0 else None
            k2 = (torch.rand(n, hidden, generator=g) >= p).to(torch.uint8) if (p > 0 and hidden) else None
            masks.append((k1, k2))
        out.append((x, y, masks))
    return out


def test_multitask_oracle_reduces_to_single_task_oracle():
    dim, c, n, p = 48, 5, 40, 0.3
    state = O.seeded_head_params(dim, c, 0, seed=3)
    batches = _batches(dim, {"a": c}, 0, n, p, 2, 1)
    cw = torch.rand(c) + 0.5
    multi, hist = probe_oracle.reference_multitask_steps({"a": state}, {"a": c}, 0, batches, p, {"a": cw}, {"a": 1.0},
                                                         lr=1e-3, weight_decay=1e-2)
    single, losses = probe_oracle.reference_train_steps(state, [(x, y[:, 0], m[0][0]) for x, y, m in batches], c, p, cw,
                                                        lr=1e-3, weight_decay=1e-2)
    for k in single:
        torch.testing.assert_close(multi["a"][k], single[k], rtol=1e-6, atol=1e-7)
    assert all(abs(h["a"] - l) < 1e-9 for h, l in zip(hist, losses))


CASES = [  # dim, tasks, hidden, batch, dropout
    (1152, {"age": 9}, 512, 256, 0.3),        # deeper single-task probe (LinearProbe(deeper_head=True))
    (1024, MIVIA, 0, 128, 0.3),               # config 3's multitask probe, linear heads
    (144, {"gender": 2, "emotion": 7}, 64, 33, 0.0),
    (1408, MIVIA, 512, 64, 0.5),
]


@pytest.mark.gpu
@pytest.mark.parametrize("dim,tasks,hidden,n,p", CASES)
def test_multitask_steps_match_reference(cuda_device, dim, tasks, hidden, n, p):
    from vlmfw_b200.probe_train import MultiTaskProbeTrainStep

    g = torch.Generator().manual_seed(dim + n)
    states = {t: O.seeded_head_params(dim, c, hidden, seed=20 + i) for i, (t, c) in enumerate(tasks.items())}
    cws = {t: (torch.rand(c, generator=g) + 0.5 if i % 2 == 0 else None) for i, (t, c) in enumerate(tasks.items())}
    tws = {t: 0.5 + 0.4 * i for i, t in enumerate(tasks)}
    batches = _batches(dim, tasks, hidden, n, p, 3, 7)
    ref, hist = probe_oracle.reference_multitask_steps(states, tasks, hidden, batches, p, cws, tws, lr=1e-3, weight_decay=1e-2)
    ts = MultiTaskProbeTrainStep(dim, tasks, "cuda", deeper=hidden > 0, hidden_dim=hidden, lr=1e-3, weight_decay=1e-2,
                                 dropout_p=p, class_weights=cws, task_weights=tws)
    ts.load_state_dict({f"heads.{t}.{k}": v.cuda() for t, s in states.items() for k, v in s.items()})
    dev = lambda m: None if m is None else m.cuda()  # noqa: E731
    for (x, y, masks), want in zip(batches, hist):
        got = ts.step(x.cuda(), y.cuda(), masks=[(dev(a), dev(b)) for a, b in masks])
        for t in tasks:
            assert abs(got[t] - want[t]) < 3e-5 * max(1.0, abs(want[t])), (t, got[t], want[t])
    sd = ts.state_dict()
    for t in tasks:
        for k, v in ref[t].items():
            if k.endswith("num_batches_tracked"):
                assert int(sd[f"heads.{t}.{k}"]) == int(v)
            else:
                # Adam normalises every gradient to a step of ~lr: an element whose true gradient is ~0 (below the fp32
                # noise of the sums) can land up to a few lr away from the fp64 oracle. Those are a vanishing fraction;
                # everything else must agree tightly.
                got, lr = sd[f"heads.{t}.{k}"].cpu().double(), 1e-3
                diff = (got - v.double()).abs()
                loose = diff > 3e-6 + 3e-4 * v.double().abs()
                assert loose.double().mean() < 1e-4 and diff.max() < 3 * lr * len(batches), \
                    f"{t}.{k}: {int(loose.sum())} of {loose.numel()} elements off, max {float(diff.max()):.3e}"


@pytest.mark.gpu
def test_deeper_sharded_phases_equal_single_process(cuda_device):
    from vlmfw_b200.parallel import shard_range
    from vlmfw_b200.probe_train import MultiTaskProbeTrainStep

    dim, tasks, hidden, n, p = 1152, {"gender": 2, "age": 9}, 512, 256, 0.3
    states = {t: O.seeded_head_params(dim, c, hidden, seed=5 + i) for i, (t, c) in enumerate(tasks.items())}
    sd = {f"heads.{t}.{k}": v.cuda() for t, s in states.items() for k, v in s.items()}
    x, y, _ = _batches(dim, tasks, hidden, n, p, 1, 3)[0]
    x, y = x.cuda(), y.cuda()

    def fresh():
        t = MultiTaskProbeTrainStep(dim, tasks, "cuda", deeper=True, hidden_dim=hidden, dropout_p=p, seed=11)
        t.load_state_dict(sd)
        return t

    single = fresh()
    masks = [single.dropout_keep_masks(i, n, 0, n) for i in range(len(tasks))]
    single.phase1_stats(x)
    full = single.phase2_grads(x, y, n, masks).clone()
    ranks = [fresh() for _ in range(4)]
    spans = [shard_range(n, r, 4) for r in range(4)]
    stats = sum(t.phase1_stats(x[lo:hi]).clone() for t, (lo, hi) in zip(ranks, spans))
    total = torch.zeros_like(full)
    for t, (lo, hi) in zip(ranks, spans):
        t.stats.copy_(stats)
        local = [t.dropout_keep_masks(i, n, lo, hi) for i in range(len(tasks))]
        assert all(torch.equal(a[0], m[0][lo:hi]) and torch.equal(a[1], m[1][lo:hi]) for a, m in zip(local, masks))
        total += t.phase2_grads(x[lo:hi], y[lo:hi], n, local)
    torch.testing.assert_close(total, full, rtol=2e-4, atol=2e-5)


@pytest.mark.gpu
def test_unlabelled_task_is_left_untouched_and_ema_weights(cuda_device):
    from vlmfw_b200.probe_train import MultiTaskProbeTrainStep

    dim, tasks = 64, {"a": 3, "b": 4}
    ts = MultiTaskProbeTrainStep(dim, tasks, "cuda", dropout_p=0.0, use_running_means=True, ema_alpha=0.9,
                                 task_weights={"a": 2.0, "b": 0.5})
    assert ts.on_epoch_start() == pytest.approx({"a": 0.4, "b": 1.6})  # 1/static, normalised to mean 1 (no EMA yet)
    before = ts.heads["b"].params.clone()
    x = torch.randn(16, dim).cuda()
    y = torch.stack([torch.randint(0, 3, (16,)), torch.full((16,), -1)], dim=1).cuda()
    losses = ts.step(x, y)
    assert losses["b"] is None and torch.equal(ts.heads["b"].params, before)
    assert not torch.equal(ts.heads["a"].params, ts.heads["a"].params * 0) and ts.ema["b"] is None
    first = losses["a"]
    y2 = torch.stack([torch.randint(0, 3, (16,)), torch.randint(0, 4, (16,))], dim=1).cuda()
    l2 = ts.step(x, y2)
    assert ts.ema["a"] == pytest.approx(0.9 * first + 0.1 * l2["a"]) and ts.ema["b"] == pytest.approx(l2["b"])
    raw = [1 / ts.ema["a"], 1 / ts.ema["b"]]
    assert ts.on_epoch_start() == pytest.approx({"a": raw[0] / (sum(raw) / 2), "b": raw[1] / (sum(raw) / 2)})


@pytest.mark.gpu
@pytest.mark.parametrize("dim,tasks,hidden,n,p", [(1024, MIVIA, 0, 128, 0.3), (1152, {"age": 9, "gender": 2}, 512, 256, 0.3),
                                                  (144, {"gender": 2, "emotion": 7}, 64, 33, 0.0)])
def test_multitask_graph_step_equals_eager_step(cuda_device, dim, tasks, hidden, n, p):
    """The multitask / deeper step replayed from one CUDA graph (device-side step counter for the dropout keys and AdamW's
    bias corrections, task weights in a device vector) = the eager step: same masks, same losses, same parameters; also
    when the task weights change between steps (running means) and when eager and graph steps alternate."""
    from vlmfw_b200.probe_train import MultiTaskProbeTrainStep

    steps = 5
    g = torch.Generator().manual_seed(3)
    xs = [(torch.randn(n, dim, generator=g) * 1.2 + 0.1).cuda() for _ in range(steps)]
    ys = [torch.stack([torch.randint(-1, c, (n,), generator=g) for c in tasks.values()], dim=1).cuda() for _ in range(steps)]

    def fresh():
        return MultiTaskProbeTrainStep(dim, tasks, "cuda", deeper=hidden > 0, hidden_dim=hidden, dropout_p=p, seed=4,
                                       use_running_means=True, ema_alpha=0.8, lr=1e-3)

    eager, graph, mixed = fresh(), fresh(), fresh()
    graph.capture(n)
    mixed.step(xs[0], ys[0])
    mixed.capture(n)
    with pytest.raises(ValueError):
        graph.step_graph(xs[0][:5], ys[0][:5])
    for s in range(steps):
        if s == 2:  # epoch boundary: the task weights move to the inverse running means
            for t in (eager, graph, mixed):
                t.on_epoch_start()
            assert graph.current_task_weights == pytest.approx(eager.current_task_weights)
        le = eager.step(xs[s], ys[s])
        lg = graph.step_graph(xs[s].cpu() if s % 2 else xs[s], ys[s])
        for t in tasks:
            assert (le[t] is None) == (lg[t] is None)
            if le[t] is not None:
                assert abs(le[t] - lg[t]) <= 2e-6 * max(1.0, abs(le[t])), (s, t, le[t], lg[t])
        if s >= 1:
            lm = mixed.step_graph(xs[s], ys[s]) if s % 2 else mixed.step(xs[s], ys[s])
            for t in tasks:
                if le[t] is not None:
                    assert abs(le[t] - lm[t]) <= 2e-6 * max(1.0, abs(le[t]))
    for other in (graph, mixed):
        assert other.step_count == eager.step_count == steps
        for t in tasks:
            torch.testing.assert_close(other.heads[t].params, eager.heads[t].params, rtol=0, atol=5e-7)
            torch.testing.assert_close(other.heads[t].running_var, eager.heads[t].running_var, rtol=1e-6, atol=1e-7)
            assert other.heads[t].num_batches_tracked == eager.heads[t].num_batches_tracked
            torch.testing.assert_close(other.last_logits[t], eager.last_logits[t], rtol=1e-5, atol=1e-6)
        other.release_graph()

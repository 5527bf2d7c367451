"""GPU parity tests of the fused RPO transformer-embedding policy forward (csrc/evac_policy.cuh, called through the
C ABI `evac_policy_forward`) -- SURVEY section 8 row f1.  Checkers: the golden produced by the UNMODIFIED reference
network (tests/golden/policy/gen_policy_golden.py) and the plain-PyTorch float32 restatement `RPOTransformerPolicy`
(itself pinned to that golden by tests/test_rollout_cpu.py).  Tolerance: float32, 2e-5 absolute on O(1) outputs
(LayerNorm'd embeddings, tanh MLP heads); the summation order differs from cuBLAS / SDPA."""
import math
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))


def _mods():
    from evacuation_b200.rollout import FusedRPOTransformerPolicy, PolicyRollout, RPOTransformerPolicy, VectorNormalizer, normalize_reward_fused
    return FusedRPOTransformerPolicy, PolicyRollout, RPOTransformerPolicy, VectorNormalizer, normalize_reward_fused


def _make(n_ped, d_model, seed=0, **kw):
    Fused, _, Torch, _, _ = _mods()
    torch.manual_seed(seed)
    net = Torch((n_ped + 2) * d_model, n_ped, **kw).cuda().eval()
    with torch.no_grad():  # non-trivial LayerNorm affine, biases and log-std (the defaults are 1 / 0 / 0)
        for name, p in net.named_parameters():
            if "norm" in name or name.endswith("bias") or name == "actor_logstd":
                p.add_(0.3 * torch.randn_like(p))
    fused = Fused(net, n_ped, device="cuda", seed=7).eval()
    return net, fused


def _check(net, fused, x, atol=2e-5):
    E = x.shape[0]
    with torch.no_grad():
        emb_t = net.embed(x)
        mean_t = net.actor_mean(emb_t)
        val_t = net.critic(emb_t).flatten()
    emb = torch.empty_like(x)
    mean, val = torch.empty((E, fused.A), device="cuda"), torch.empty(E, device="cuda")
    fused.forward(x, embedding=emb, mean=mean, value=val, sample=False)
    torch.cuda.synchronize()
    np.testing.assert_allclose(emb.cpu().numpy(), emb_t.cpu().numpy(), rtol=0, atol=atol)
    np.testing.assert_allclose(mean.cpu().numpy(), mean_t.cpu().numpy(), rtol=0, atol=atol)
    np.testing.assert_allclose(val.cpu().numpy(), val_t.cpu().numpy(), rtol=0, atol=atol * 4)


def test_fused_policy_matches_reference_network_golden():
    """Outputs of the unmodified reference RPOTransformerEmbedding (eval mode) on its own inputs."""
    Fused, _, Torch, _, _ = _mods()
    z = np.load(os.path.join(HERE, "golden", "policy", "policy_transformer.npz"))
    n = int(z["number_of_pedestrians"])
    x = torch.as_tensor(z["x"]).cuda()
    net = Torch(x.shape[1], n).eval()
    net.load_state_dict({k[2:]: torch.as_tensor(z[k]) for k in z.files if k.startswith("w:")}, strict=True)
    fused = Fused(net, n, device="cuda").eval()
    E = x.shape[0]
    emb = torch.empty_like(x)
    o = {k: torch.empty(s, device="cuda") for k, s in (("mean", (E, 2)), ("value", (E,)), ("logprob", (E,)), ("entropy", (E,)), ("action", (E, 2)))}
    fused.forward(x, embedding=emb, given_action=torch.as_tensor(z["sampled_action"]).cuda(), **o)
    torch.cuda.synchronize()
    np.testing.assert_allclose(emb.cpu().numpy(), z["embedding"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(o["mean"].cpu().numpy(), z["actor_mean"], rtol=0, atol=2e-5)
    np.testing.assert_allclose(o["value"].cpu().numpy(), z["value"].reshape(-1), rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(o["logprob"].cpu().numpy(), z["logprob_of_sampled"], rtol=1e-5, atol=5e-5)
    np.testing.assert_allclose(o["entropy"].cpu().numpy(), z["entropy"], rtol=1e-6)
    np.testing.assert_array_equal(o["action"].cpu().numpy(), z["sampled_action"])


def test_fused_policy_matches_reference_network_variants():
    """Hyper-parameter variants evaluated by the UNMODIFIED reference module (gen_policy_variants_golden.py): residual
    blocks, 1 / 2 / 4 heads, d_model 3 / 2, 1 / 3 blocks, narrow feed-forward, 62 and 64 rows."""
    Fused, _, Torch, _, _ = _mods()
    z = np.load(os.path.join(HERE, "golden", "policy", "policy_variants.npz"))
    for i in range(int(z["num_variants"])):
        pre = f"v{i}:"
        n, d, heads, dff, blocks, resid, nh = (int(v) for v in z[pre + "cfg"])
        net = Torch((n + 2) * d, n, num_heads=heads, dim_feedforward=dff, num_blocks=blocks, use_resid=bool(resid), num_hidden=nh).eval()
        net.load_state_dict({k[len(pre) + 2:]: torch.as_tensor(z[k]) for k in z.files if k.startswith(pre + "w:")}, strict=True)
        fused = Fused(net, n, device="cuda").eval()
        x = torch.as_tensor(z[pre + "x"]).cuda()
        E = x.shape[0]
        emb = torch.empty_like(x)
        o = {k: torch.empty(s, device="cuda") for k, s in (("mean", (E, 2)), ("value", (E,)), ("logprob", (E,)))}
        fused.forward(x, embedding=emb, given_action=torch.as_tensor(z[pre + "action"]).cuda(), **o)
        torch.cuda.synchronize()
        tol = 6e-5 if resid else 2e-5
        np.testing.assert_allclose(emb.cpu().numpy(), z[pre + "embedding"], rtol=0, atol=tol, err_msg=f"variant {i}")
        np.testing.assert_allclose(o["mean"].cpu().numpy(), z[pre + "actor_mean"], rtol=0, atol=tol, err_msg=f"variant {i}")
        np.testing.assert_allclose(o["value"].cpu().numpy(), z[pre + "value"].reshape(-1), rtol=1e-5, atol=4 * tol, err_msg=f"variant {i}")
        np.testing.assert_allclose(o["logprob"].cpu().numpy(), z[pre + "logprob"], rtol=1e-5, atol=1e-4, err_msg=f"variant {i}")


@pytest.mark.parametrize("E", [1, 3, 70, 1000])
def test_fused_policy_default_shape_vs_torch(E):
    net, fused = _make(60, 6, seed=E)
    x = (torch.randn(E, 372, device="cuda") * 0.7).clamp_(-1, 1)
    _check(net, fused, x)


@pytest.mark.parametrize("n_ped,d_model,kw", [
    (60, 3, {}),                                  # statuses = cat
    (60, 2, {}),                                  # statuses = no
    (5, 6, {}), (30, 6, {}), (31, 6, {}), (62, 6, {}), (1, 6, {}),   # S = 7, 32, 33, 64 (the limit), 3
    (17, 3, {}), (9, 2, {}),
    (60, 6, dict(num_heads=1)), (60, 6, dict(num_heads=2)), (60, 6, dict(num_heads=4)),
    (60, 6, dict(dim_feedforward=10)), (60, 6, dict(dim_feedforward=33, num_blocks=3)), (60, 6, dict(num_blocks=1)),
    (60, 6, dict(use_resid=True)), (60, 6, dict(num_hidden=32)), (20, 6, dict(num_hidden=12, use_resid=True)),
])
def test_fused_policy_shapes_vs_torch(n_ped, d_model, kw):
    net, fused = _make(n_ped, d_model, seed=3, **kw)
    x = (torch.randn(37, (n_ped + 2) * d_model, device="cuda") * 0.7).clamp_(-1, 1)
    _check(net, fused, x, atol=4e-5 if kw.get("use_resid") else 2e-5)


@pytest.mark.parametrize("E", [5, 129, 4096, 38000])
def test_heads_on_tensor_cores_match_the_cuda_core_path_and_float64(E, monkeypatch):
    """The heads ([E x 372] x [372 x 128], then 64 x 64 per head) run as 3xTF32 tcgen05 MMAs (csrc/evac_policy_tc.cuh; 64-column tiles
    below 296 row tiles, 128-column tiles above: E = 38000).  EVAC_POLICY_TC=0 (read at evac_policy_create) keeps it on the
    CUDA cores.  Both must sit within float32 rounding of a float64 evaluation of the heads on the kernel's own embedding
    [rpo_linear_agent_network.py:23-42]."""
    Fused, _, Torch, _, _ = _mods()
    torch.manual_seed(11)
    net = Torch(372, 60).cuda().eval()
    with torch.no_grad():
        for name, p in net.named_parameters():
            if name.endswith("bias"):
                p.add_(0.3 * torch.randn_like(p))
    x = (torch.randn(E, 372, device="cuda") * 0.7).clamp_(-1, 1)
    outs = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("EVAC_POLICY_TC", tc)
        fused = Fused(net, 60, device="cuda", seed=7).eval()
        emb = torch.empty_like(x)
        mean, val, lp = torch.empty((E, 2), device="cuda"), torch.empty(E, device="cuda"), torch.empty(E, device="cuda")
        act = torch.empty((E, 2), device="cuda")
        fused.forward(x, embedding=emb, mean=mean, value=val, action=act, logprob=lp, sample=True)
        torch.cuda.synchronize()
        outs[tc] = (emb, mean, val, act, lp)
    assert torch.equal(outs["1"][0], outs["0"][0])
    net64 = Torch(372, 60).double().cuda().eval()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    with torch.no_grad():
        e64 = outs["1"][0].double()
        mean64, val64 = net64.actor_mean(e64), net64.critic(e64).flatten()
    for tc in ("1", "0"):
        assert float((outs[tc][1].double() - mean64).abs().max()) < 2e-6, tc
        assert float((outs[tc][2].double() - val64).abs().max() / val64.abs().max().clamp_min(1.0)) < 6e-6, tc
    # the sampled action is mean + std * noise with noise keyed by (seed, call, env): identical streams on both paths
    np.testing.assert_allclose(outs["1"][3].cpu().numpy(), outs["0"][3].cpu().numpy(), rtol=0, atol=2e-6)
    np.testing.assert_allclose(outs["1"][4].cpu().numpy(), outs["0"][4].cpu().numpy(), rtol=0, atol=2e-5)


def test_fused_policy_large_logits_take_the_exact_softmax_path():
    """Attention logits of magnitude ~1e2-1e3: the Cauchy-Schwarz shift underflows whole rows, which must be redone
    with the exact maximum (near one-hot softmax); compared with a float64 evaluation of the torch restatement."""
    net, fused = _make(60, 6, seed=11)
    with torch.no_grad():
        for blk in net.embedding:
            blk.attention.Wq.weight.mul_(40.0); blk.attention.Wk.weight.mul_(40.0)
    fused.load_from(net)
    x = (torch.randn(64, 372, device="cuda")).clamp_(-1, 1)
    with torch.no_grad():
        emb32 = net.embed(x)
        emb64 = net.double().embed(x.double())
    emb = torch.empty_like(x)
    fused.forward(x, embedding=emb)
    torch.cuda.synchronize()
    assert torch.isfinite(emb).all()
    # near-ties of huge logits are ill-conditioned in float32 for ANY implementation: the fused kernel must be about as
    # close to the float64 evaluation as the float32 torch forward is
    err, err32 = (emb.double() - emb64).abs(), (emb32.double() - emb64).abs()
    assert float(err.median()) <= 5 * float(err32.median()) + 1e-6, (float(err.median()), float(err32.median()))
    assert float(err.mean()) <= 5 * float(err32.mean()) + 1e-5, (float(err.mean()), float(err32.mean()))


def test_fused_policy_nan_rows_propagate():
    net, fused = _make(60, 6, seed=2)
    x = (torch.randn(8, 372, device="cuda") * 0.5).clamp_(-1, 1)
    x[3, 17] = float("nan")
    emb = torch.empty_like(x)
    fused.forward(x, embedding=emb)
    with torch.no_grad():
        ref = net.embed(x)
    assert torch.isnan(emb[3]).any() and torch.equal(torch.isnan(emb).any(dim=1), torch.isnan(ref).any(dim=1))
    np.testing.assert_allclose(emb[:3].cpu().numpy(), ref[:3].cpu().numpy(), rtol=0, atol=2e-5)


def test_fused_normalizer_prologue_matches_vector_normalizer():
    """NormalizeObservation + clip fused into the policy kernel == VectorNormalizer.observation (float32, same update
    order), and the reward normaliser kernel == VectorNormalizer.reward, over a sequence of steps."""
    Fused, _, Torch, Norm, norm_reward = _mods()
    net, fused = _make(60, 6, seed=5)
    E, D = 50, 372
    n_ref, n_fused = Norm(E, D, device="cuda"), Norm(E, D, device="cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    for t in range(12):
        raw = torch.randn(E, D, device="cuda", generator=g) * (1 + t) + 0.5
        want = n_ref.observation(raw)
        got, emb = torch.empty_like(raw), torch.empty_like(raw)
        fused.forward(raw, embedding=emb, normalizer=n_fused, obs_norm=got)
        np.testing.assert_allclose(got.cpu().numpy(), want.cpu().numpy(), rtol=0, atol=1e-6)
        np.testing.assert_allclose(n_fused.obs_mean.cpu().numpy(), n_ref.obs_mean.cpu().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(n_fused.obs_var.cpu().numpy(), n_ref.obs_var.cpu().numpy(), rtol=1e-6, atol=1e-7)
        assert float(n_fused.obs_count) == float(n_ref.obs_count)
        with torch.no_grad():
            np.testing.assert_allclose(emb.cpu().numpy(), net.embed(want).cpu().numpy(), rtol=0, atol=3e-5)
        rew = torch.randn(E, device="cuda", generator=g) * 3 - 1
        term = torch.rand(E, device="cuda", generator=g) < 0.2
        want_r = n_ref.reward(rew, term)
        trunc = torch.rand(E, device="cuda", generator=g) < 0.1
        done = torch.empty(E, device="cuda")
        got_r = norm_reward(n_fused, rew, term, torch.empty(E, device="cuda"), truncated=trunc, done_out=done)
        assert torch.equal(done, (term | trunc).float())
        np.testing.assert_allclose(got_r.cpu().numpy(), want_r.cpu().numpy(), rtol=1e-6, atol=1e-7)
        np.testing.assert_allclose(n_fused.returns.cpu().numpy(), n_ref.returns.cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_fused_policy_sampling_statistics_and_logprob():
    net, fused = _make(60, 6, seed=9)
    E = 20000
    x = (torch.randn(E, 372, device="cuda") * 0.7).clamp_(-1, 1)
    mean, act, clip, lp, ent, val = (torch.empty(s, device="cuda") for s in ((E, 2), (E, 2), (E, 2), (E,), (E,), (E,)))
    fused.forward(x, mean=mean, action=act, action_clipped=clip, logprob=lp, entropy=ent, value=val)
    std = torch.exp(net.actor_logstd.detach()).expand_as(mean)
    zs = ((act - mean) / std).cpu().numpy()
    assert abs(zs.mean()) < 0.03 and abs(zs.std() - 1.0) < 0.03 and abs(np.corrcoef(zs[:, 0], zs[:, 1])[0, 1]) < 0.03
    assert abs((np.abs(zs) > 1.96).mean() - 0.05) < 0.01
    dist = torch.distributions.Normal(mean, std)
    np.testing.assert_allclose(lp.cpu().numpy(), dist.log_prob(act).sum(1).cpu().numpy(), rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(ent.cpu().numpy(), dist.entropy().sum(1).cpu().numpy(), rtol=1e-6)
    assert torch.equal(clip, act.clamp(-1, 1))
    # the device-side call counter advances the stream: a second call draws different numbers, same means
    act2, mean2 = torch.empty_like(act), torch.empty_like(mean)
    fused.forward(x, mean=mean2, action=act2)
    assert torch.equal(mean, mean2) and not torch.equal(act, act2)
    # same (seed, offset, global env index) -> same numbers, whatever the batch split (sharding invariance)
    Fused = _mods()[0]
    fa = Fused(net, 60, device="cuda", seed=7).eval()
    fb = Fused(net, 60, device="cuda", seed=7, env_index_offset=100).eval()
    a_full, a_part = torch.empty((E, 2), device="cuda"), torch.empty((50, 2), device="cuda")
    fa.forward(x, action=a_full)
    fb.forward(x[100:150].contiguous(), action=a_part)
    assert torch.equal(a_full[100:150], a_part)


def test_fused_policy_dropout_mode():
    """training=True applies dropout (reference rollouts never call .eval()): p = 0 is the eval forward bit for bit;
    p = 0.1 changes the embedding, is deterministic per (seed, offset) and differs between calls."""
    Fused, _, Torch, _, _ = _mods()
    torch.manual_seed(0)
    x = (torch.randn(256, 372, device="cuda") * 0.7).clamp_(-1, 1)
    net0 = Torch(372, 60, dropout=0.0).cuda()
    f0 = Fused(net0, 60, device="cuda")
    a, b = torch.empty_like(x), torch.empty_like(x)
    f0.train().forward(x, embedding=a)
    f0.eval().forward(x, embedding=b)
    assert torch.equal(a, b)
    net1 = Torch(372, 60, dropout=0.1).cuda()
    net1.load_state_dict(net0.state_dict())
    f1 = Fused(net1, 60, device="cuda", seed=3)
    c, d, e = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    f1.train().forward(x, embedding=c, advance=False)
    f1.forward(x, embedding=d)          # same offset -> identical masks
    f1.forward(x, embedding=e)          # advanced offset -> new masks
    assert torch.equal(c, d) and not torch.equal(c, e) and not torch.equal(c, b)
    assert torch.isfinite(c).all() and torch.isfinite(e).all()
    # dropout is a perturbation, not a different function: typical deviation from the eval output is O(1) but bounded,
    # and the eval mode of the same handle still equals the p = 0 network
    f1.eval().forward(x, embedding=d)
    np.testing.assert_allclose(d.cpu().numpy(), b.cpu().numpy(), rtol=0, atol=1e-6)
    frac_changed = float(((c - b).abs() > 1e-4).float().mean())
    assert 0.5 < frac_changed <= 1.0


def test_policy_rollout_fused_graph_runs_and_matches_torch_loop_in_eval_mode():
    """Config-5 loop: fused policy (normaliser prologue, heads, ClipAction) -> fused env step -> reward normaliser,
    CUDA-graph replayed.  With dropout off and sampling replaced by the mean, the fused loop retraces the PyTorch loop."""
    import evacuation_b200 as eb
    Fused, Rollout, Torch, Norm, norm_reward = _mods()
    E, N = 96, 60
    mk = lambda: eb.setup_env(eb.EnvConfig(number_of_pedestrians=N, is_new_exiting_reward=True), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                              num_envs=E, seed=5, auto_reset=True)
    torch.manual_seed(1)
    net = Torch(372, N).cuda()
    # (1) the production loop: graph-captured, dropout on, sampling on
    ro = Rollout(mk(), Fused(net, N, device="cuda", seed=1), use_graph=True, store=True)
    ro.reset()
    buf = ro.run(20)
    torch.cuda.synchronize()
    for k in ("obs", "actions", "logprobs", "rewards", "values", "dones"):
        assert torch.isfinite(buf[k]).all(), k
    assert float(buf["obs"].abs().max()) <= 1.0 and not torch.equal(buf["actions"][0], buf["actions"][1])
    # (2) eval-mode retrace against torch: same env seed, actions = clipped means
    env_a, env_b = mk(), mk()
    fused = Fused(net, N, device="cuda").eval()
    net.eval()
    na, nb = Norm(E, 372, device="cuda"), Norm(E, 372, device="cuda")
    oa, _ = env_a.reset(); ob, _ = env_b.reset()
    oa, ob = oa.reshape(E, 372), ob.reshape(E, 372)
    mean, val, clip, xn = torch.empty((E, 2), device="cuda"), torch.empty(E, device="cuda"), torch.empty((E, 2), device="cuda"), torch.empty((E, 372), device="cuda")
    for t in range(15):
        fused.forward(oa, normalizer=na, obs_norm=xn, mean=mean, value=val, action_clipped=clip, sample=False)
        with torch.no_grad():
            xb = nb.observation(ob)
            mb = net.actor_mean(net.embed(xb))
        np.testing.assert_allclose(xn.cpu().numpy(), xb.cpu().numpy(), rtol=0, atol=2e-6)
        np.testing.assert_allclose(mean.cpu().numpy(), mb.cpu().numpy(), rtol=0, atol=2e-5)
        act = mb.clamp(-1, 1).contiguous()  # both envs take the SAME action so that the traces stay comparable
        oa2, ra, ta, _, _ = env_a.step(act)
        ob2, rb, tb, _, _ = env_b.step(act)
        assert torch.equal(oa2, ob2) and torch.equal(ra, rb)
        oa, ob = oa2.reshape(E, 372), ob2.reshape(E, 372)
        got_r = norm_reward(na, ra, ta, torch.empty(E, device="cuda"))
        np.testing.assert_allclose(got_r.cpu().numpy(), nb.reward(rb, tb).cpu().numpy(), rtol=1e-6, atol=1e-7)


def test_unsupported_policy_shapes_raise():
    Fused, _, Torch, _, _ = _mods()
    with pytest.raises(NotImplementedError):
        Fused(Torch(70 * 6, 68), 68, device="cuda")          # 70 rows > one warp's 64
    with pytest.raises(NotImplementedError):
        Fused(Torch(62 * 6, 60, num_hidden=128), 60, device="cuda")
    with pytest.raises(NotImplementedError):
        Fused(Torch(62 * 6, 60, num_heads=5), 60, device="cuda")


def test_rollout_capture_does_not_advance_the_environments():
    """the warm-up iterations before the CUDA-graph capture used to step the real environments (episodes started
    two transitions in, normaliser counts off by two).  They run on a snapshot now: after reset() the first run(T) must
    record exactly T transitions from the reset state -- same as the eager (no graph) loop, bit for bit in eval mode."""
    import evacuation_b200 as eb
    from evacuation_b200.rollout import FusedRPOTransformerPolicy, PolicyRollout, RPOTransformerPolicy

    E, n, T = 6, 60, 5
    outs = []
    for use_graph in (True, False):
        env = eb.setup_env(eb.EnvConfig(number_of_pedestrians=n), eb.EnvWrappersConfig(positions="rel", statuses="ohe", type="Box"),
                           num_envs=E, seed=4, auto_reset=True)
        torch.manual_seed(3)
        net = RPOTransformerPolicy(env.unwrapped.obs_dim, n).cuda()
        pol = FusedRPOTransformerPolicy(net, n, device="cuda", seed=9).eval()
        ro = PolicyRollout(env, pol, use_graph=use_graph, store=True)
        ro.reset()
        buf = ro.run(T)
        st = env.unwrapped.get_state()
        assert bool((st["now"] == T).all()), st["now"]
        assert int(pol.calls) == T and abs(float(ro.norm.obs_count) - (1e-4 + T)) < 1e-9 and abs(float(ro.norm.ret_count) - (1e-4 + T)) < 1e-9
        assert bool((buf["dones"][0] == 0).all())
        outs.append((buf, st))
    for k in outs[0][0]:
        assert torch.equal(outs[0][0][k], outs[1][0][k]), k
    for k in outs[0][1]:
        assert torch.equal(outs[0][1][k], outs[1][1][k]), k

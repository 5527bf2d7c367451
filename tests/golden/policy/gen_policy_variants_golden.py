"""Generate tests/golden/policy/policy_variants.npz: outputs of the UNMODIFIED reference network
(/root/reference/src/agents/networks/rpo_transformer_agent_network.py) for hyper-parameter variants (residual blocks,
head counts, d_model 3 / 2, block counts, feed-forward widths, the full 60-pedestrian shape), eval() mode, fixed weights
and inputs.  Same import shim as gen_policy_golden.py; run in the authoring container only:
    python tests/golden/policy/gen_policy_variants_golden.py
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

REF = "/root/reference/src/agents/networks"
pkg = types.ModuleType("refnets")
pkg.__path__ = [REF]
sys.modules["refnets"] = pkg
for name in ("utils", "rpo_linear_agent_network", "rpo_transformer_agent_network"):
    spec = importlib.util.spec_from_file_location(f"refnets.{name}", os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[f"refnets.{name}"] = mod
    spec.loader.exec_module(mod)
lin, tr = sys.modules["refnets.rpo_linear_agent_network"], sys.modules["refnets.rpo_transformer_agent_network"]

# (number_of_pedestrians, d_model, kwargs of RPOTransformerEmbeddingConfig, num_hidden, batch)
VARIANTS = [
    (10, 6, dict(use_resid=True), 16, 4),
    (7, 3, dict(num_heads=2), 16, 4),
    (12, 2, dict(num_heads=4, dim_feedforward=40, num_blocks=1), 16, 4),
    (30, 6, dict(num_heads=1, num_blocks=3, dim_feedforward=10), 32, 3),
    (60, 6, dict(), 16, 3),            # the BASELINE shape (62 rows), narrow heads to keep the fixture small
    (62, 6, dict(use_resid=True), 8, 2),   # 64 rows: the one-warp limit of the fused kernel
]


class _Space:
    def __init__(self, shape):
        self.shape = shape


out = {"num_variants": np.array(len(VARIANTS))}
for i, (n, d, kw, nh, batch) in enumerate(VARIANTS):
    obs_dim = (n + 2) * d

    class _Envs:
        single_observation_space = _Space((obs_dim,))
        single_action_space = _Space((2,))

    torch.manual_seed(100 + i)
    cfg = tr.RPOTransformerEmbeddingConfig(network=lin.RPOLinearNetworkConfig(num_hidden=nh), **kw)
    net = tr.RPOTransformerEmbedding(_Envs(), n, cfg, torch.device("cpu")).eval()
    with torch.no_grad():  # the defaults leave LayerNorm affine / biases / log-std trivial (1 / 0 / 0)
        for name, p in net.named_parameters():
            if "norm" in name or name.endswith("bias") or name == "actor_logstd":
                p.add_(0.3 * torch.randn_like(p))
    x = (torch.randn(batch, obs_dim) * 0.8).clamp_(-1, 1)
    act = torch.randn(batch, 2)
    with torch.no_grad():
        emb = net.embedding(x)
        mean = net.actor_mean(emb)
        value = net.get_value(x)
        std = torch.exp(net.actor_logstd.expand_as(mean))
        logprob = torch.distributions.Normal(mean, std).log_prob(act).sum(1)
    pre = f"v{i}:"
    out[pre + "cfg"] = np.array([n, d, cfg.num_heads, int(cfg.dim_feedforward), cfg.num_blocks, int(cfg.use_resid), nh])
    for k, v in (("x", x), ("embedding", emb), ("actor_mean", mean), ("value", value), ("action", act), ("logprob", logprob)):
        out[pre + k] = v.numpy()
    for k, v in net.state_dict().items():
        out[pre + "w:" + k] = v.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_variants.npz"), **out)
print(len(VARIANTS), "variants", sum(v.nbytes for v in out.values()) // 1024, "KiB raw")

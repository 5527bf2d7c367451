"""Generate tests/golden/policy_transformer.npz from the UNMODIFIED reference network
(/root/reference/src/agents/networks/rpo_transformer_agent_network.py), imported here as a bare
package (its parent `agents/__init__` pulls gymnasium, which is not installed).  eval() mode
(dropout off), fixed weights and inputs; run in the authoring container only:
    python tests/golden/gen_policy_golden.py
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

N, D = 10, 6
obs_dim = (N + 2) * D


class _Space:
    def __init__(self, shape):
        self.shape = shape


class _Envs:
    single_observation_space = _Space((obs_dim,))
    single_action_space = _Space((2,))


torch.manual_seed(7)
cfg = tr.RPOTransformerEmbeddingConfig(network=lin.RPOLinearNetworkConfig())
net = tr.RPOTransformerEmbedding(_Envs(), N, cfg, torch.device("cpu")).eval()
x = torch.randn(5, obs_dim)
act = torch.randn(5, 2)
with torch.no_grad():
    emb = net.embedding(x)
    mean = net.actor_mean(emb)
    value = net.get_value(x)
    torch.manual_seed(11)
    a2, logprob, entropy, v2 = net.get_action_and_value(x)
out = {"x": x.numpy(), "embedding": emb.numpy(), "actor_mean": mean.numpy(), "value": value.numpy(), "sampled_action": a2.numpy(),
       "logprob_of_sampled": logprob.numpy(), "entropy": entropy.numpy(), "number_of_pedestrians": np.array(N)}
for k, v in net.state_dict().items():
    out["w:" + k] = v.numpy()
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "policy_transformer.npz"), **out)
print({k: v.shape for k, v in out.items() if not k.startswith("w:")}, len([k for k in out if k.startswith("w:")]), "weight tensors")

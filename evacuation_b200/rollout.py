"""Rollout-loop glue around the fused step kernel (BASELINE config 5).

The caller of the hot path in the reference is the cleanrl-style rollout loop of
`src/agents/rpo_agent.py:180-203`: policy forward -> `envs.step(action)` -> per-env gymnasium
wrappers (`wrapping()`, rpo_agent.py:24-33: FlattenObservation, ClipAction, NormalizeObservation,
clip(-1, 1), NormalizeReward(gamma), clip(-100, 100)).  Everything here is the device-resident,
batched equivalent of that glue so that the loop never leaves the GPU:

* `VectorNormalizer`   -- per-environment running mean / variance of observations and of the discounted
  return (gymnasium `RunningMeanStd`, batch of one sample per step and env), with the two clips;
* `RPOTransformerPolicy` -- the RPO transformer-embedding actor-critic
  (`src/agents/networks/rpo_transformer_agent_network.py:36-163`,
  `rpo_linear_agent_network.py:19-61`) restated in plain PyTorch (library GEMMs / SDPA: this is
  the consumer of the step kernel, not a kernel target of this tier);
* `PolicyRollout`      -- policy -> step -> normalise, one iteration captured in a CUDA graph
  and replayed (the tiny-GEMM policy is otherwise launch-bound).

The environment step is the CUDA path; there is no CPU fallback here either.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.nn as nn
import torch.nn.functional as F


# ---------------------------------------------------------------------------------------------
class VectorNormalizer:
    """NormalizeObservation + clip(-1,1) + NormalizeReward(gamma) + clip(-100,100), one independent
    running estimate per environment (the reference wraps every env before vectorising them).

    gymnasium's `RunningMeanStd.update` with a batch of one sample x (batch_var = 0, batch_count = 1):
        delta = x - mean;  tot = count + 1
        mean' = mean + delta / tot
        var'  = (var * count + delta^2 * count / tot) / tot
    starting from mean = 0, var = 1, count = 1e-4; normalised value = (x - mean') / sqrt(var' + 1e-8).
    """

    def __init__(self, num_envs: int, obs_dim: int, gamma: float = 0.99, device="cuda", epsilon: float = 1e-8,
                 obs_clip: float = 1.0, reward_clip: float = 100.0, dtype=torch.float32):
        self.gamma, self.epsilon, self.obs_clip, self.reward_clip = gamma, epsilon, obs_clip, reward_clip
        self.obs_mean = torch.zeros((num_envs, obs_dim), dtype=dtype, device=device)
        self.obs_var = torch.ones((num_envs, obs_dim), dtype=dtype, device=device)
        self.ret_mean = torch.zeros(num_envs, dtype=dtype, device=device)
        self.ret_var = torch.ones(num_envs, dtype=dtype, device=device)
        self.returns = torch.zeros(num_envs, dtype=dtype, device=device)
        # the observation and return estimators see exactly one sample per call, so they share one count
        # per stream; kept as device scalars so that a captured CUDA graph advances them
        self.obs_count = torch.full((), 1e-4, dtype=torch.float64, device=device)
        self.ret_count = torch.full((), 1e-4, dtype=torch.float64, device=device)

    @staticmethod
    def _update(mean, var, count, x):
        tot = count + 1.0
        w_new = (1.0 / tot).to(mean.dtype)
        w_old = (count / tot).to(mean.dtype)
        delta = x - mean
        mean.add_(delta * w_new)
        var.mul_(w_old).add_(delta * delta * (w_old * w_new))
        count.add_(1.0)

    def observation(self, obs: torch.Tensor) -> torch.Tensor:
        """obs [E, D] (flattened like FlattenObservation) -> normalised, clipped copy."""
        self._update(self.obs_mean, self.obs_var, self.obs_count, obs)
        out = (obs - self.obs_mean) * torch.rsqrt(self.obs_var + self.epsilon)
        return out.clamp_(-self.obs_clip, self.obs_clip)

    def reward(self, reward: torch.Tensor, terminated: torch.Tensor) -> torch.Tensor:
        """NormalizeReward.step: returns = returns * gamma * (1 - terminated) + reward; reward / sqrt(var(returns) + eps)."""
        self.returns.mul_(self.gamma * (1.0 - terminated.to(self.returns.dtype))).add_(reward)
        self._update(self.ret_mean, self.ret_var, self.ret_count, self.returns)
        out = reward * torch.rsqrt(self.ret_var + self.epsilon)
        return out.clamp_(-self.reward_clip, self.reward_clip)


# ---------------------------------------------------------------------------------------------
def _ortho(layer: nn.Linear, std: float = math.sqrt(2.0), bias: float = 0.0) -> nn.Linear:
    nn.init.orthogonal_(layer.weight, std)  # networks/utils.py:4-7
    nn.init.constant_(layer.bias, bias)
    return layer


class _SetAttention(nn.Module):
    """The reference's attention variant (rpo_transformer_agent_network.py:36-75): projections to
    d_model * num_heads, then the softmax runs over the sequence with the d_model axis acting as the
    batch-of-heads axis and num_heads as the contracted feature axis; scale = sqrt(d_model)."""

    def __init__(self, d_model: int, num_heads: int):
        super().__init__()
        self.d_model, self.num_heads = d_model, num_heads
        self.Wq = nn.Linear(d_model, d_model * num_heads)
        self.Wk = nn.Linear(d_model, d_model * num_heads)
        self.Wv = nn.Linear(d_model, d_model * num_heads)
        self.dense = nn.Linear(d_model * num_heads, d_model)

    def forward(self, x):  # [B, S, D]
        B, S, D = x.shape
        H = self.num_heads

        def split(t):  # [B, S, H*D] -> [B, D, S, H]
            return t.view(B, S, H, D).permute(0, 3, 1, 2)

        # one GEMM for the three projections (K = 6 GEMMs are launch / tail bound); parameter names stay the reference's
        qkv = F.linear(x, torch.cat([self.Wq.weight, self.Wk.weight, self.Wv.weight]), torch.cat([self.Wq.bias, self.Wk.bias, self.Wv.bias]))
        q, k, v = (split(t) for t in qkv.split(H * D, dim=-1))
        if x.is_cuda:
            # zero-padding the contracted axis (H = 3) to 8 changes neither the scores nor the outputs but lets SDPA take
            # its fused kernel instead of the batched-GEMM + softmax fallback (2.6 -> 1.1 ms at 8192 envs on B200)
            pad = (-H) % 8
            q, k, v = (F.pad(t, (0, pad)) for t in (q, k, v))
            out = F.scaled_dot_product_attention(q, k, v, scale=1.0 / math.sqrt(D))[..., :H]
        else:
            out = F.scaled_dot_product_attention(q, k, v, scale=1.0 / math.sqrt(D))  # [B, D, S, H]
        out = out.transpose(1, 2).reshape(B, S, D * H)
        return self.dense(out)


def _layer_norm(x, ln: nn.LayerNorm):
    """LayerNorm over a 6-wide last axis written as elementwise ops: torch's row-wise kernel launches one block per
    row and takes 2.2 ms for the 5e5 rows of an 8192-env batch (half of the whole policy forward on B200)."""
    mu = x.mean(dim=-1, keepdim=True)
    xc = x - mu
    var = (xc * xc).mean(dim=-1, keepdim=True)
    return xc * torch.rsqrt(var + ln.eps) * ln.weight + ln.bias


class _Block(nn.Module):
    """rpo_transformer_agent_network.py:78-131: attention -> dropout -> (resid) -> LayerNorm -> FF (Linear, Dropout,
    ReLU, Linear) -> dropout -> (resid) -> LayerNorm."""

    def __init__(self, d_model: int, num_heads: int, d_ff: int, dropout: float, use_resid: bool):
        super().__init__()
        self.d_model, self.use_resid = d_model, use_resid
        self.attention = _SetAttention(d_model, num_heads)
        self.ff = nn.Sequential(nn.Linear(d_model, d_ff), nn.Dropout(dropout), nn.ReLU(), nn.Linear(d_ff, d_model))
        self.norm1, self.norm2 = nn.LayerNorm(d_model), nn.LayerNorm(d_model)
        self.dropout = nn.Dropout(dropout)

    def forward(self, x):
        shape = x.shape
        if x.dim() == 2:
            x = x.view(shape[0], -1, self.d_model)
        a = self.dropout(self.attention(x))
        x = _layer_norm(x + a if self.use_resid else a, self.norm1)
        f = self.dropout(self.ff(x))
        x = _layer_norm(x + f if self.use_resid else f, self.norm2)
        return x.view(shape)


class RPOTransformerPolicy(nn.Module):
    """RPOTransformerEmbedding (rpo_transformer_agent_network.py:133-163) over RPOLinearNetwork
    (rpo_linear_agent_network.py:19-61).  `obs_dim` = (N + 2) * d_model flattened MatrixObs row."""

    def __init__(self, obs_dim: int, number_of_pedestrians: int, action_dim: int = 2, num_hidden: int = 64, rpo_alpha: float = 0.5,
                 num_blocks: int = 2, num_heads: int = 3, dim_feedforward: int = 96, dropout: float = 0.1, use_resid: bool = False,
                 chunk: int = 8192):
        super().__init__()
        d_model = obs_dim // (number_of_pedestrians + 2)
        self.rpo_alpha, self.chunk = rpo_alpha, chunk
        self.embedding = nn.Sequential(*[_Block(d_model, num_heads, dim_feedforward, dropout, use_resid) for _ in range(num_blocks)])
        self.critic = nn.Sequential(_ortho(nn.Linear(obs_dim, num_hidden)), nn.Tanh(), _ortho(nn.Linear(num_hidden, num_hidden)), nn.Tanh(),
                                    _ortho(nn.Linear(num_hidden, 1), std=1.0))
        self.actor_mean = nn.Sequential(_ortho(nn.Linear(obs_dim, num_hidden)), nn.Tanh(), _ortho(nn.Linear(num_hidden, num_hidden)), nn.Tanh(),
                                        _ortho(nn.Linear(num_hidden, action_dim), std=0.01))
        self.actor_logstd = nn.Parameter(torch.zeros(1, action_dim))

    def embed(self, x):
        if x.shape[0] <= self.chunk:
            return self.embedding(x)
        return torch.cat([self.embedding(c) for c in x.split(self.chunk)], dim=0)  # bounds the attention workspace

    def get_value(self, x):
        return self.critic(self.embed(x))

    def get_action_and_value(self, x, action: Optional[torch.Tensor] = None):
        x = self.embed(x)
        mean = self.actor_mean(x)
        std = torch.exp(self.actor_logstd.expand_as(mean))
        if action is None:
            action = mean + std * torch.randn_like(mean)  # Normal(mean, std).sample()
        else:  # RPO: perturb the mean when re-evaluating stored actions
            mean = mean + torch.empty_like(mean).uniform_(-self.rpo_alpha, self.rpo_alpha)
        var = std * std
        logprob = (-((action - mean) ** 2) / (2 * var) - torch.log(std) - 0.5 * math.log(2 * math.pi)).sum(1)
        entropy = (0.5 + 0.5 * math.log(2 * math.pi) + torch.log(std)).sum(1)
        return action, logprob, entropy, self.critic(x)


# ---------------------------------------------------------------------------------------------
class PolicyRollout:
    """policy forward -> ClipAction -> fused env step (same-step auto-reset) -> normalise, `num_steps` times.

    One iteration is captured in a CUDA graph (`use_graph=True`) and replayed; storage tensors follow
    rpo_agent.py:160-163 (obs, actions, logprobs, rewards, dones, values)."""

    def __init__(self, env, policy: RPOTransformerPolicy, gamma: float = 0.99, use_graph: bool = True, store: bool = True):
        self.env, self.policy, self.store, self.use_graph = env, policy, store, use_graph
        u = env.unwrapped
        self.E, self.D, self.device = u.num_envs, u.obs_dim, u.device
        self.norm = VectorNormalizer(self.E, self.D, gamma=gamma, device=self.device)
        self.next_obs = torch.zeros((self.E, self.D), device=self.device)
        self.next_done = torch.zeros(self.E, device=self.device)
        self.out = {k: torch.zeros(s, device=self.device) for k, s in
                    (("action", (self.E, 2)), ("logprob", (self.E,)), ("value", (self.E,)), ("reward", (self.E,)))}
        self._graph = None

    def reset(self):
        obs, _ = self.env.reset()
        self.next_obs.copy_(self.norm.observation(obs.reshape(self.E, self.D)))
        self.next_done.zero_()

    @torch.no_grad()
    def _iteration(self):
        action, logprob, _, value = self.policy.get_action_and_value(self.next_obs)
        obs, reward, term, trunc, _ = self.env.step(action.clamp(-1.0, 1.0).contiguous())  # ClipAction
        self.out["action"].copy_(action); self.out["logprob"].copy_(logprob); self.out["value"].copy_(value.flatten())
        self.out["reward"].copy_(self.norm.reward(reward, term))
        self.next_obs.copy_(self.norm.observation(obs.reshape(self.E, self.D)))
        self.next_done.copy_(torch.logical_or(term, trunc).to(self.next_done.dtype))

    def _capture(self):
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):
            for _ in range(2):  # warm-up outside the capture (lazy initialisation, autotuning)
                self._iteration()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph):
            self._iteration()

    def run(self, num_steps: int):
        """Returns the storage dict ([T, E, ...]) when `store`, else None."""
        if self.use_graph and self._graph is None:
            self._capture()
        buf = None
        if self.store:
            buf = dict(obs=torch.empty((num_steps, self.E, self.D), device=self.device), actions=torch.empty((num_steps, self.E, 2), device=self.device),
                       logprobs=torch.empty((num_steps, self.E), device=self.device), rewards=torch.empty((num_steps, self.E), device=self.device),
                       dones=torch.empty((num_steps, self.E), device=self.device), values=torch.empty((num_steps, self.E), device=self.device))
        for t in range(num_steps):
            if buf is not None:
                buf["obs"][t].copy_(self.next_obs); buf["dones"][t].copy_(self.next_done)
            if self._graph is not None:
                self._graph.replay()
            else:
                self._iteration()
            if buf is not None:
                buf["actions"][t].copy_(self.out["action"]); buf["logprobs"][t].copy_(self.out["logprob"])
                buf["values"][t].copy_(self.out["value"]); buf["rewards"][t].copy_(self.out["reward"])
        return buf


__all__ = ["VectorNormalizer", "RPOTransformerPolicy", "PolicyRollout"]

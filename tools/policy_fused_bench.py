"""Fused policy forward (evac_policy_forward) alone: embedding kernel and heads kernel timed separately, eval and
training (dropout) mode, against the PyTorch restatement.  python tools/policy_fused_bench.py [E]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from evacuation_b200.rollout import FusedRPOTransformerPolicy, RPOTransformerPolicy, VectorNormalizer

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
net = RPOTransformerPolicy(372, 60).cuda()
fused = FusedRPOTransformerPolicy(net, 60, device="cuda", max_envs=E)
x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)
emb = torch.empty_like(x)
mean, val, act, clip, lp = (torch.empty(s, device="cuda") for s in ((E, 2), (E,), (E, 2), (E, 2), (E,)))
norm = VectorNormalizer(E, 372, device="cuda")
xn = torch.empty_like(x)


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


out = {"E": E}
fused.eval()
out["embed_eval_us"] = timeit(lambda: fused.forward(x, embedding=emb))
out["embed+heads_eval_us"] = timeit(lambda: fused.forward(x, embedding=emb, mean=mean, value=val, action=act, action_clipped=clip, logprob=lp))
fused.train()
out["embed_train_us"] = timeit(lambda: fused.forward(x, embedding=emb))
out["full_train_norm_us"] = timeit(lambda: fused.forward(x, normalizer=norm, obs_norm=xn, mean=mean, value=val, action=act, action_clipped=clip, logprob=lp))
with torch.no_grad():
    out["torch_train_us"] = timeit(lambda: net.get_action_and_value(x), n=5)
# algorithmic flops of the forward per env (multiply-add = 2): projections, scores, weighted sums, dense, ff, heads
S, D, H, F, NH = 62, 6, 3, 96, 64
per_block = 2 * (S * D * 3 * H * D + 2 * D * S * S * H + S * D * H * D + 2 * S * D * F)
flops = 2 * per_block + 2 * (2 * (S * D * NH + NH * NH) + 3 * NH)
out["flops_per_env"] = flops
out["tflops_full_train"] = flops * E / out["full_train_norm_us"] / 1e6
print(json.dumps(out))

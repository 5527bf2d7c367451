"""Layer 1 of the policy heads: tensor-core kernel (tcgen05, 3xTF32) vs the CUDA-core path (EVAC_POLICY_TC=0), same weights and
inputs: output differences against a float64 evaluation of the PyTorch network, and the time of the heads part (forward
with heads minus embedding only).  python tools/policy_tc_ab.py [E ...]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from evacuation_b200.rollout import FusedRPOTransformerPolicy, RPOTransformerPolicy


def timeit(fn, n=50):
    for _ in range(5):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


for E in [int(v) for v in sys.argv[1:]] or [8192, 65536]:
    torch.manual_seed(0)
    net = RPOTransformerPolicy(372, 60).cuda().eval()
    x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)
    out = {"E": E}
    res = {}
    for tc in ("1", "0"):
        os.environ["EVAC_POLICY_TC"] = tc
        fused = FusedRPOTransformerPolicy(net, 60, device="cuda", max_envs=E)
        fused.eval()
        emb = torch.empty_like(x)
        mean, val, act, lp = (torch.empty(s, device="cuda") for s in ((E, 2), (E,), (E, 2), (E,)))
        fused.forward(x, embedding=emb, mean=mean, value=val, action=act, logprob=lp, sample=False)
        torch.cuda.synchronize()
        res[tc] = (emb.clone(), mean.clone(), val.clone())
        t_e = timeit(lambda: fused.forward(x, embedding=emb))
        t_f = timeit(lambda: fused.forward(x, embedding=emb, mean=mean, value=val, action=act, logprob=lp, sample=False))
        out[f"tc{tc}_embed_us"] = round(t_e, 2)
        out[f"tc{tc}_embed+heads_us"] = round(t_f, 2)
        out[f"tc{tc}_heads_us"] = round(t_f - t_e, 2)
    # float64 heads on the kernel's own embedding
    net64 = RPOTransformerPolicy(372, 60).double().cuda().eval()
    net64.load_state_dict({k: v.double() for k, v in net.state_dict().items()})
    with torch.no_grad():
        e64 = res["1"][0].double()
        ref_mean, ref_val = net64.actor_mean(e64), net64.critic(e64).squeeze(-1)
    for tc in ("1", "0"):
        out[f"tc{tc}_mean_err"] = float((res[tc][1].double() - ref_mean).abs().max())
        out[f"tc{tc}_value_err"] = float((res[tc][2].double() - ref_val).abs().max())
    out["emb_identical"] = bool(torch.equal(res["1"][0], res["0"][0]))
    print(json.dumps(out), flush=True)

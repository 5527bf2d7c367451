import os, sys, json
sys.path.insert(0, os.getcwd())
import torch
from evacuation_b200.rollout import FusedRPOTransformerPolicy, RPOTransformerPolicy
def timeit(fn, n=30):
    for _ in range(5): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for E in [int(v) for v in sys.argv[1:]] or [8192, 65536, 1001]:
    torch.manual_seed(0)
    net = RPOTransformerPolicy(372, 60).cuda()
    x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)
    out = {"E": E}; res = {}
    for mode in ("warp", "half"):
        os.environ["EVAC_POLICY_EMBED"] = mode
        fused = FusedRPOTransformerPolicy(net, 60, device="cuda", max_envs=E, seed=5)
        emb = torch.empty_like(x)
        fused.eval(); fused.forward(x, embedding=emb); torch.cuda.synchronize(); res[mode + "_eval"] = emb.clone()
        out[mode + "_eval_us"] = round(timeit(lambda: fused.forward(x, embedding=emb)), 2)
        fused.train(); fused.forward(x, embedding=emb, advance=False); torch.cuda.synchronize(); res[mode + "_train"] = emb.clone()
        out[mode + "_train_us"] = round(timeit(lambda: fused.forward(x, embedding=emb)), 2)
    out["eval_identical"] = bool(torch.equal(res["warp_eval"], res["half_eval"]))
    out["train_identical"] = bool(torch.equal(res["warp_train"], res["half_train"]))
    out["eval_maxdiff"] = float((res["warp_eval"] - res["half_eval"]).abs().max())
    print(json.dumps(out), flush=True)

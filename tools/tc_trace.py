import os, sys
sys.path.insert(0, os.getcwd())
import torch
from evacuation_b200.rollout import FusedRPOTransformerPolicy, RPOTransformerPolicy
torch.manual_seed(0)
net = RPOTransformerPolicy(372, 60).cuda().eval()
for E in (1000, 8192):
    x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)
    fused = FusedRPOTransformerPolicy(net, 60, device="cuda", max_envs=E).eval()
    mean, val = torch.empty((E, 2), device="cuda"), torch.empty(E, device="cuda")
    emb = torch.empty_like(x)
    for i in range(3):
        fused.forward(x, embedding=emb, mean=mean, value=val, sample=False)
    torch.cuda.synchronize()
    os.environ["EVAC_TC_TRACE_PRINT"] = "1"
    print("E", E, file=sys.stderr)
    for i in range(3):
        fused.forward(x, embedding=emb, mean=mean, value=val, sample=False)
    del os.environ["EVAC_TC_TRACE_PRINT"]

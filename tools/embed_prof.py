import os, sys
sys.path.insert(0, os.getcwd())
import torch
from evacuation_b200.rollout import FusedRPOTransformerPolicy, RPOTransformerPolicy
torch.manual_seed(0)
E = 8192
net = RPOTransformerPolicy(372, 60).cuda()
fused = FusedRPOTransformerPolicy(net, 60, device="cuda", max_envs=E)
fused.train()
x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)
emb = torch.empty_like(x)
for i in range(4):
    fused.forward(x, embedding=emb)
torch.cuda.synchronize()

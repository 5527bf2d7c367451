"""Where does the PyTorch policy forward spend its time (config 5)?  python tools/policy_bench.py [E]"""
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

from evacuation_b200.rollout import RPOTransformerPolicy

E = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
torch.manual_seed(0)
net = RPOTransformerPolicy(372, 60).cuda()
x = torch.randn(E, 372, device="cuda").clamp_(-1, 1)


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


with torch.no_grad():
    print("full get_action_and_value ms:", timeit(lambda: net.get_action_and_value(x)))
    print("embed only ms:", timeit(lambda: net.embed(x)))
    blk = net.embedding[0]
    xb = x.view(E, 62, 6)
    att = blk.attention
    print("attention ms:", timeit(lambda: att(xb)))
    q = att.Wq(xb).view(E, 62, 3, 6).permute(0, 3, 1, 2)
    print("sdpa (head dim 3) ms:", timeit(lambda: F.scaled_dot_product_attention(q, q, q, scale=1 / math.sqrt(6))))
    for pad in (8, 16):
        qp = F.pad(q, (0, pad - 3)).contiguous()
        print(f"sdpa padded to {pad} fp32 ms:", timeit(lambda: F.scaled_dot_product_attention(qp, qp, qp, scale=1 / math.sqrt(6))))
        qh = qp.to(torch.bfloat16)
        print(f"sdpa padded to {pad} bf16 ms:", timeit(lambda: F.scaled_dot_product_attention(qh, qh, qh, scale=1 / math.sqrt(6))))
    print("ff ms:", timeit(lambda: blk.ff(xb)))
    print("layernorm ms:", timeit(lambda: blk.norm1(xb)))
    print("dropout ms:", timeit(lambda: blk.dropout(xb)))
    emb = net.embed(x)
    print("actor+critic mlp ms:", timeit(lambda: (net.actor_mean(emb), net.critic(emb))))
    from torch.profiler import ProfilerActivity, profile

    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        net.get_action_and_value(x)
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=14, max_name_column_width=70))

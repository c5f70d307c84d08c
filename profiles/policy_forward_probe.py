"""Where the rollout step goes once the engine is 1 % of it: the MF-Q policy forward on one
group's observation block [E * cap, 13, 13, 7] as the engine leaves it in HBM, timed in the variants that keep the
network PyTorch.     python profiles/policy_forward_probe.py [rows]"""
import sys
import time

import torch
import torch.nn.functional as F

sys.path.insert(0, "mean-field-multi-agent-reinforcement-learning_b200/python")
from mfmarl_b200.algo.base import QNet  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
dev = torch.device("cuda")
torch.manual_seed(0)
net = QNet((13, 13, 7), (34,), 21, use_mf=True).to(dev).eval()
view = (torch.rand(N, 13, 13, 7, device=dev) < 0.1).float() * torch.rand(N, 13, 13, 7, device=dev)
feat, prob = torch.rand(N, 34, device=dev), torch.softmax(torch.rand(N, 21, device=dev), 1)


def timed(name, fn, reps=5):
    with torch.no_grad():
        for _ in range(2):
            out = fn()
        torch.cuda.synchronize()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(reps):
            out = fn()
        t1.record()
        torch.cuda.synchronize()
    ms = t0.elapsed_time(t1) / reps
    print("%-58s %8.3f ms  %6.2f Mrows/s  peak mem %.1f GB" % (name, ms, N / ms / 1e3, torch.cuda.max_memory_allocated() / 1e9), flush=True)
    return out


ref = timed("fp32 (as shipped)", lambda: net(view, feat, prob))
torch.backends.cuda.matmul.allow_tf32 = True
torch.backends.cudnn.allow_tf32 = True
tf32 = timed("tf32", lambda: net(view, feat, prob))


def autocast_fn():
    with torch.autocast("cuda", dtype=torch.bfloat16):
        return net(view, feat, prob)


bf = timed("bf16 autocast (act_autocast)", autocast_fn)

net_cl = QNet((13, 13, 7), (34,), 21, use_mf=True).to(dev).eval()
net_cl.load_state_dict(net.state_dict())
net_cl = net_cl.to(memory_format=torch.channels_last)
timed("tf32 + channels_last weights", lambda: net_cl(view, feat, prob))
net_h = QNet((13, 13, 7), (34,), 21, use_mf=True).to(dev).eval()
net_h.load_state_dict(net.state_dict())
net_h = net_h.to(torch.bfloat16).to(memory_format=torch.channels_last)
cl = timed("bf16 weights + channels_last, inputs cast per call", lambda: net_h(view.to(torch.bfloat16), feat.to(torch.bfloat16), prob.to(torch.bfloat16)))


def split_fn():     # the same in 4 slices: does the working set matter?
    return torch.cat([net_h(view[i::4].to(torch.bfloat16), feat[i::4].to(torch.bfloat16), prob[i::4].to(torch.bfloat16)) for i in range(4)])


timed("  ... in 4 slices", split_fn)

# the first conv as ONE matmul over im2col patches of the 7-channel view padded to 8 channels
w1 = torch.zeros(32, 3, 3, 8, device=dev, dtype=torch.bfloat16)
w1[..., :7] = net.conv1.weight.permute(0, 2, 3, 1).to(torch.bfloat16)


def conv1_matmul():
    x = F.pad(view.to(torch.bfloat16), (0, 1))                               # [N, 13, 13, 8]
    p = x.unfold(1, 3, 1).unfold(2, 3, 1)                                     # [N, 11, 11, 8, 3, 3]
    p = p.permute(0, 1, 2, 4, 5, 3).reshape(N * 121, 72)
    return torch.relu(p @ w1.reshape(32, 72).t() + net.conv1.bias.to(torch.bfloat16))


timed("conv1 alone as im2col matmul (bf16)", conv1_matmul)
timed("conv1 alone, cudnn bf16 channels_last", lambda: torch.relu(net_h.conv1(view.to(torch.bfloat16).permute(0, 3, 1, 2))))
timed("conv1+conv2, cudnn bf16 channels_last", lambda: torch.relu(net_h.conv2(torch.relu(net_h.conv1(view.to(torch.bfloat16).permute(0, 3, 1, 2))))))
agree = lambda a: float((a.float().argmax(1) == ref.argmax(1)).float().mean())
print("argmax agreement with fp32: tf32 %.4f  bf16 autocast %.4f  bf16 channels_last %.4f" % (agree(tf32), agree(bf), agree(cl)))
with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
    with torch.no_grad():
        net_h(view.to(torch.bfloat16), feat.to(torch.bfloat16), prob.to(torch.bfloat16))
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))

"""ValueNet: the Q-network shared by MF-Q and IL (reference algo/base.py:7-281, TF1 graph) in PyTorch.

Architecture (base.py:123-190): view -> conv3x3x32 -> conv3x3x32 -> dense 256; feature -> dense 32;
[mean action -> dense 64 -> dense 32 when use_mf]; concat -> dense 128 -> dense 64 -> dense n_actions.
Eval and target copies, soft target update tau = 0.005 (base.py:84-92), masked MSE on the taken action
(base.py:94-116), Adam 1e-4, gamma = 0.95.  Inputs may be numpy arrays (single-env binding) or CUDA tensors
(batched engine: observations and mean actions never leave the device).
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


def as_tensor(x, device, dtype=torch.float32):
    if isinstance(x, torch.Tensor):
        return x.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.ascontiguousarray(x), dtype=dtype).to(device, non_blocking=True)


def checkpoint_path(dir_path, stem, step):
    """<dir>/<stem>_<step>.pt -- the reference writes tf.train.Saver files "<stem>_<step>" per scope."""
    return os.path.join(dir_path, "%s_%s.pt" % (stem, step))


def sync_gradients(parameters):
    """Shared-parameter data-parallel training (north star: "NCCL over NVLink is used only for the optional
    shared-parameter gradient allreduce"): average the gradients of `parameters` over the ranks of the default
    process group with ONE all-reduce of a flat buffer (1-2 MB for these nets: latency-bound, NVLS does the sum).
    A no-op when torch.distributed is not initialised or the world has one rank."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return
    flat = torch.cat([g.reshape(-1) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat.div_(dist.get_world_size())
    offset = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[offset:offset + n].view_as(g))
        offset += n


def agree_on_count(n, device):
    """Every rank must run the same number of synchronised optimiser steps: take the minimum over the ranks."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return n
    t = torch.tensor([n], dtype=torch.int64, device=device if dist.get_backend() == "nccl" else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t[0])


def _dense(n_in, n_out):
    layer = nn.Linear(n_in, n_out)
    nn.init.xavier_uniform_(layer.weight)      # tf.layers default: glorot_uniform, zero bias
    nn.init.zeros_(layer.bias)
    return layer


def _conv_bias_relu(conv, x, fused):
    """relu(conv(x) + bias).  `fused` (the bf16 rollout twin on CUDA, no autograd): cuDNN's convolution with the bias and
    the ReLU in its epilogue -- one pass over the output instead of three (PyTorch's conv adds the bias with a
    broadcasting elementwise kernel and the ReLU is one more; at 65 536 rows those two passes were 1.3 ms of the 2.9 ms
    MF-Q forward, profiles/r02/play_busy_probe.txt)."""
    if fused and x.is_cuda and not torch.is_grad_enabled():
        try:
            return torch.cudnn_convolution_relu(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation,
                                                conv.groups)
        except RuntimeError:
            pass
    return F.relu(conv(x))


class QNet(nn.Module):
    fused_conv = False        # set on the bf16 rollout twin (bf16_rollout_copy)

    def __init__(self, view_space, feature_space, num_actions, use_mf):
        super().__init__()
        h, w, c = view_space
        self.use_mf = use_mf
        self.conv1 = nn.Conv2d(c, 32, 3)
        self.conv2 = nn.Conv2d(32, 32, 3)
        for conv in (self.conv1, self.conv2):
            nn.init.xavier_uniform_(conv.weight)
            nn.init.zeros_(conv.bias)
        self.dense_obs = _dense(32 * (h - 4) * (w - 4), 256)
        self.dense_emb = _dense(feature_space[0], 32)
        width = 256 + 32
        if use_mf:
            self.prob_emb = _dense(num_actions, 64)
            self.dense_act_prob = _dense(64, 32)
            width += 32
        self.dense2 = _dense(width, 128)
        self.dense_out = _dense(128, 64)
        self.q_value = _dense(64, num_actions)

    def forward(self, view, feature, prob=None):
        x = view.permute(0, 3, 1, 2)                      # engine layout is NHWC: a channels_last view, no copy
        x = _conv_bias_relu(self.conv2, _conv_bias_relu(self.conv1, x, self.fused_conv), self.fused_conv)
        x = x.permute(0, 2, 3, 1).flatten(1)              # flatten in (H, W, C) order like the TF graph (base.py:134-136)
        parts = [F.relu(self.dense_obs(x)), F.relu(self.dense_emb(feature))]
        if self.use_mf:
            parts.append(F.relu(self.dense_act_prob(F.relu(self.prob_emb(prob)))))
        x = F.relu(self.dense2(torch.cat(parts, dim=1)))
        return self.q_value(F.relu(self.dense_out(x)))


def load_tf_variables(net, layers, variables):
    """Import TensorFlow-layout weights into a torch module: the correspondence that makes these networks the
    reference's graphs, stated in code (and what a user with a TF checkpoint of the reference needs).

    layers     [(tf_layer_name, torch_attribute)], e.g. TF_QNET_LAYERS
    variables  {"<tf_layer_name>/kernel": array, "<tf_layer_name>/bias": array} as tf.train.load_checkpoint returns
               them: conv kernels HWIO over an NHWC input (tf.layers.conv2d, base.py:125-136), dense kernels [in, out]
               applied as x @ kernel + bias (tf.layers.dense).  A flattened conv output is in (h, w, c) order in both
               graphs (QNet.forward flattens NHWC), so dense kernels need nothing but the transpose.
    Layers the module does not have (the mean-action branch without use_mf) are skipped; missing arrays raise."""
    with torch.no_grad():
        for tf_name, attr in layers:
            layer = getattr(net, attr, None)
            if layer is None:
                continue
            kernel = torch.as_tensor(np.asarray(variables[tf_name + "/kernel"]), dtype=torch.float32)
            bias = torch.as_tensor(np.asarray(variables[tf_name + "/bias"]), dtype=torch.float32)
            kernel = kernel.permute(3, 2, 0, 1) if kernel.dim() == 4 else kernel.t()      # HWIO -> OIHW; [in, out] -> [out, in]
            assert tuple(kernel.shape) == tuple(layer.weight.shape), (tf_name, tuple(kernel.shape), tuple(layer.weight.shape))
            layer.weight.copy_(kernel.contiguous())
            layer.bias.copy_(bias)


# variable scopes of ValueNet._construct_net (base.py:123-183) -> QNet attributes
TF_QNET_LAYERS = [("Conv1", "conv1"), ("Conv2", "conv2"), ("Dense-Obs", "dense_obs"), ("Dense-Emb", "dense_emb"),
                  ("Prob-Emb", "prob_emb"), ("Dense-Act-Prob", "dense_act_prob"), ("Dense2", "dense2"),
                  ("Dense-Out", "dense_out"), ("Q-Value", "q_value")]


def _pad_last_input_channel(weight, h, w, c):
    """dense weight [out, h*w*c] over a flattened (h, w, c) view -> [out, h*w*(c+1)] with zeros for the extra channel"""
    out = weight.new_zeros((weight.shape[0], h * w, c + 1))
    out[:, :, :c] = weight.reshape(weight.shape[0], h * w, c)
    return out.reshape(weight.shape[0], -1)


def bf16_rollout_copy(net):
    """A bf16, channels-last copy of a QNet for the ROLLOUT forward on the engine's bf16 observation rows
    (BatchedGridWorld.observe_groups(dtype=torch.bfloat16): [N, 13, 13, 8], channel 7 = 0): conv1 takes the eighth
    channel with zero weights, so no cast and no padding pass stand between k_obs and the tensor cores.  Training
    always uses the fp32 network; refresh the copy after the weights change."""
    import copy
    twin = copy.deepcopy(net)
    c_in = net.conv1.in_channels
    if c_in % 8:
        conv1 = nn.Conv2d(c_in + 1, net.conv1.out_channels, net.conv1.kernel_size)
        with torch.no_grad():
            conv1.weight.zero_()
            conv1.weight[:, :c_in] = net.conv1.weight
            conv1.bias.copy_(net.conv1.bias)
        twin.conv1 = conv1.to(net.conv1.weight.device)
    twin.fused_conv = True
    return twin.to(torch.bfloat16).to(memory_format=torch.channels_last).eval()


class ValueNet:
    def __init__(self, env, handle, name, update_every=5, use_mf=False, learning_rate=1e-4, tau=0.005, gamma=0.95,
                 device=None):
        self.env, self.name, self.handle = env, name, handle
        self.view_space = tuple(env.get_view_space(handle))
        assert len(self.view_space) == 3
        self.feature_space = tuple(env.get_feature_space(handle))
        self.num_actions = env.get_action_space(handle)[0]
        self.update_every, self.use_mf = update_every, use_mf
        self.temperature = 0.1
        self.grad_sync = False        # True: average gradients over the torch.distributed ranks before every step
        self.act_autocast = None      # e.g. torch.bfloat16: run the ROLLOUT forward (act) under autocast; training stays fp32
        self.lr, self.tau, self.gamma = learning_rate, tau, gamma
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.eval_net = QNet(self.view_space, self.feature_space, self.num_actions, use_mf).to(self.device)
        self.target_net = QNet(self.view_space, self.feature_space, self.num_actions, use_mf).to(self.device)
        self.optimizer = torch.optim.Adam(self.eval_net.parameters(), lr=learning_rate)
        self._rollout16, self._rollout16_stale = None, True     # bf16 twin of eval_net for bf16 observation rows

    # -- the reference exposes the variable list for the self-play soft copy (base.py:185-190, tools.py:566-569)
    @property
    def vars(self):
        return list(self.eval_net.parameters()) + list(self.target_net.parameters())

    def _inputs(self, view, feature, prob):
        v, f = as_tensor(view, self.device), as_tensor(feature, self.device)
        p = None
        if self.use_mf:
            assert prob is not None
            p = as_tensor(prob, self.device)
            if p.shape[0] == 1 and v.shape[0] != 1:
                p = p.expand(v.shape[0], -1)
        return v, f, p

    @torch.no_grad()
    def calc_target_q(self, **kwargs):
        """base.py:192-220: r + (1 - done) * gamma * Q_target(s', m')[argmax_a Q_eval(s', m')]."""
        v, f, p = self._inputs(kwargs["obs"], kwargs["feature"], kwargs.get("prob"))
        t_q, e_q = self.target_net(v, f, p), self.eval_net(v, f, p)
        q = t_q.gather(1, e_q.argmax(dim=1, keepdim=True)).squeeze(1)
        rewards = as_tensor(kwargs["rewards"], self.device)
        dones = as_tensor(kwargs["dones"], self.device)
        return rewards + (1.0 - dones) * q * self.gamma

    @torch.no_grad()
    def update(self):
        """soft target update (base.py:84-92, 223-226)."""
        for t, e in zip(self.target_net.parameters(), self.eval_net.parameters()):
            t.mul_(1.0 - self.tau).add_(e, alpha=self.tau)

    @torch.no_grad()
    def act(self, **kwargs):
        """argmax of softmax(Q / temperature).  As in the reference (base.py:39,82,240,253) the temperature is
        baked at 0.1 and argmax is taken, so `eps` does not change the choice.  Returns int32 actions: numpy
        for numpy inputs, a device tensor for device inputs."""
        view, feature = kwargs["state"][0], kwargs["state"][1]
        self.temperature = kwargs.get("eps", self.temperature)
        if isinstance(view, torch.Tensor) and view.dtype == torch.bfloat16:
            # the engine's bf16 rows ([N, 13, 13, 8]): straight into the bf16 channels-last twin
            if self._rollout16 is None or self._rollout16_stale:
                self._rollout16, self._rollout16_stale = bf16_rollout_copy(self.eval_net), False
            prob = kwargs.get("prob")
            p = None
            if self.use_mf:
                p = prob.to(torch.bfloat16)
                if p.shape[0] == 1 and view.shape[0] != 1:
                    p = p.expand(view.shape[0], -1)
            q = self._rollout16(view, feature.to(torch.bfloat16), p)
            return q.argmax(dim=1).to(torch.int32)
        v, f, p = self._inputs(view, feature, kwargs.get("prob"))
        if self.use_mf and not isinstance(kwargs["prob"], torch.Tensor):
            assert len(kwargs["prob"]) == len(view)
        if self.act_autocast is not None and v.is_cuda:
            with torch.autocast("cuda", dtype=self.act_autocast):
                q = self.eval_net(v, f, p)
        else:
            q = self.eval_net(v, f, p)
        actions = q.argmax(dim=1).to(torch.int32)
        return actions if isinstance(view, torch.Tensor) else actions.cpu().numpy()

    def train(self, **kwargs):
        """base.py:256-281: masked squared error between target_q and Q_eval(s, m)[a]."""
        v, f, p = self._inputs(kwargs["state"][0], kwargs["state"][1], kwargs.get("prob"))
        target_q = as_tensor(kwargs["target_q"], self.device)
        mask = as_tensor(kwargs["masks"], self.device)
        acts = as_tensor(kwargs["acts"], self.device, torch.int64)
        e_q = self.eval_net(v, f, p).gather(1, acts.unsqueeze(1)).squeeze(1)
        loss = ((target_q - e_q) ** 2 * mask).sum() / mask.sum()
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.grad_sync:
            sync_gradients(self.eval_net.parameters())
        self.optimizer.step()
        self._rollout16_stale = True
        return float(loss.detach()), {"Eval-Q": round(float(e_q.detach().mean()), 6),
                                      "Target-Q": round(float(target_q.mean()), 6)}

    # -- checkpoints (q_learning.py:53-71,146-171 use tf.train.Saver per scope)
    def _save(self, dir_path, stem, step):
        os.makedirs(dir_path, exist_ok=True)
        path = checkpoint_path(dir_path, stem, step)
        torch.save({"eval": self.eval_net.state_dict(), "target": self.target_net.state_dict(),
                    "optimizer": self.optimizer.state_dict()}, path)
        print("[*] Model saved at: {}".format(path))

    def _load(self, dir_path, stem, step):
        path = checkpoint_path(dir_path, stem, step)
        blob = torch.load(path, map_location=self.device)
        self.eval_net.load_state_dict(blob["eval"])
        self.target_net.load_state_dict(blob["target"])
        self.optimizer.load_state_dict(blob["optimizer"])
        self._rollout16_stale = True
        print("[*] Loaded model from {}".format(path))

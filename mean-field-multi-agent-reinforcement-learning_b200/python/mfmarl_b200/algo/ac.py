"""Actor-critic and mean-field actor-critic (reference algo/ac.py:8-176 `ActorCritic`, :178-361 `MFAC`) in PyTorch.

Network (ac.py:52-80 / :226-262): flatten(view) -> dense 256, feature -> dense 256, concat -> dense 512 (relu);
policy = softmax(dense(h / 0.1)) clipped to [1e-10, 1 - 1e-10], sampled with a multinomial draw;
AC value = dense(h, 1); MFAC value = dense(relu(dense([h_view, h_emb, relu(dense(relu(dense(mean action, 64)), 32))], 256)), 1)
-- the mean action only feeds the critic.  Loss (ac.py:82-95): -mean(stop_grad(R - V) * log pi(a)) +
value_coef * mean((R - V)^2) + ent_coef * mean(sum pi log pi), Adam 1e-4 (the un-clipped `minimize` op is the one
the reference keeps, :97-104).  `train` turns every agent's episode into discounted returns bootstrapped from the
value of its LAST state (ac.py:139-148), then does one update on the whole batch.
"""
import os

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import tools
from .base import _dense, as_tensor, checkpoint_path, sync_gradients
from .replay_device import DeviceEpisodesBuffer, segmented_discounted_returns


class ACNet(nn.Module):
    def __init__(self, view_space, feature_space, num_actions, use_mf):
        super().__init__()
        hidden = 256
        self.use_mf = use_mf
        self.view_dense = _dense(int(np.prod(view_space)), hidden)
        self.emb_dense = _dense(feature_space[0], hidden)
        self.trunk = _dense(2 * hidden, 2 * hidden)
        self.policy_head = _dense(2 * hidden, num_actions)
        if use_mf:
            self.prob_emb = _dense(num_actions, 64)
            self.prob_dense = _dense(64, 32)
            self.value_dense = _dense(2 * hidden + 32, hidden)
            self.value_head = _dense(hidden, 1)
        else:
            self.value_head = _dense(2 * hidden, 1)

    def _features(self, view, feature):
        both = torch.cat([F.relu(self.view_dense(view.flatten(1))), F.relu(self.emb_dense(feature))], dim=1)
        return both, F.relu(self.trunk(both))

    def policy(self, view, feature):
        _, h = self._features(view, feature)
        return torch.softmax(self.policy_head(h / 0.1), dim=1).clamp(1e-10, 1.0 - 1e-10)

    def bf16_rollout_copy(self, view_space):
        """bf16 twin for the engine's bf16 observation rows ([N, 13, 13, 8], channel 7 = 0): the first dense layer gets
        zero weights for the extra channel (see base.bf16_rollout_copy)."""
        import copy
        from .base import _pad_last_input_channel
        twin = copy.deepcopy(self)
        h, w, c = view_space
        if c % 8:
            dense = nn.Linear(h * w * (c + 1), self.view_dense.out_features).to(self.view_dense.weight.device)
            with torch.no_grad():
                dense.weight.copy_(_pad_last_input_channel(self.view_dense.weight, h, w, c))
                dense.bias.copy_(self.view_dense.bias)
            twin.view_dense = dense
        return twin.to(torch.bfloat16).eval()

    def forward(self, view, feature, prob=None):
        both, h = self._features(view, feature)
        policy = torch.softmax(self.policy_head(h / 0.1), dim=1).clamp(1e-10, 1.0 - 1e-10)
        if self.use_mf:
            p = F.relu(self.prob_dense(F.relu(self.prob_emb(prob))))
            value = self.value_head(F.relu(self.value_dense(torch.cat([both, p], dim=1))))
        else:
            value = self.value_head(h)
        return policy, value.reshape(-1)


# tf.layers.dense default names in creation order (ac.py:50-103 ActorCritic, :220-262 MFAC) -> ACNet attributes,
# for base.load_tf_variables
TF_AC_LAYERS = [("dense", "view_dense"), ("dense_1", "emb_dense"), ("dense_2", "trunk"), ("dense_3", "policy_head"),
                ("dense_4", "value_head")]
TF_MFAC_LAYERS = [("dense", "view_dense"), ("dense_1", "emb_dense"), ("dense_2", "trunk"), ("dense_3", "policy_head"),
                  ("dense_4", "prob_emb"), ("dense_5", "prob_dense"), ("dense_6", "value_dense"), ("dense_7", "value_head")]


def discounted_returns(rewards, bootstrap, gamma):
    """ac.py:144-148: keep = V(last state); for i reversed: keep = keep * gamma + r[i]; r[i] = keep."""
    out = np.array(rewards, dtype=np.float64 if np.asarray(rewards).dtype == np.float64 else np.float32)
    keep = bootstrap
    for i in range(len(out) - 1, -1, -1):
        keep = keep * gamma + out[i]
        out[i] = keep
    return out


class _ActorCriticBase:
    use_mf = False
    stem = "ac"

    def __init__(self, name, handle, env, value_coef=0.1, ent_coef=0.08, gamma=0.95, batch_size=64, learning_rate=1e-4,
                 device=None, seed=None, stage_rows=1 << 18, sub_len=400):
        self.env, self.name = env, name
        self.view_space = tuple(env.get_view_space(handle))
        self.feature_space = tuple(env.get_feature_space(handle))
        self.num_actions = env.get_action_space(handle)[0]
        self.gamma = self.reward_decay = gamma
        self.batch_size, self.learning_rate = batch_size, learning_rate
        self.value_coef, self.ent_coef = value_coef, ent_coef
        self.device = torch.device(device if device is not None else ("cuda" if torch.cuda.is_available() else "cpu"))
        self.net = ACNet(self.view_space, self.feature_space, self.num_actions, self.use_mf).to(self.device)
        self.optimizer = torch.optim.Adam(self.net.parameters(), lr=learning_rate)
        self.replay_buffer = tools.EpisodesBuffer(use_mean=self.use_mf)
        self.device_replay = None
        self.grad_sync = False        # True: average gradients over the torch.distributed ranks before every step
        self.act_autocast = None      # e.g. torch.bfloat16: run the ROLLOUT forward (act) under autocast; training stays fp32
        self._stage_cfg = (stage_rows, sub_len)
        self.generator = torch.Generator(device=self.device)
        if seed is not None:
            self.generator.manual_seed(seed)

    @property
    def vars(self):
        return list(self.net.parameters())

    def flush_buffer(self, **kwargs):
        self.replay_buffer.push(**kwargs)

    def flush_buffer_batched(self, **kwargs):
        """One lockstep step of the batched engine (device tensors [E, cap, ...]) -> device episode store."""
        if self.device_replay is None:
            self.device_replay = DeviceEpisodesBuffer(self.view_space, self.feature_space, self.num_actions,
                                                      self._stage_cfg[0], self._stage_cfg[1], use_mean=self.use_mf,
                                                      device=self.device)
        self.device_replay.push(**kwargs)

    def _train_batched(self, verbose):
        ep = self.device_replay.episodes()
        if ep is None:
            return None
        with torch.no_grad():        # bootstrap = V(last state) per trajectory (ac.py:139-143)
            last = ep["seg_last"]
            boot = self.net(ep["view"][last], ep["feature"][last], ep["prob"][last] if self.use_mf else None)[1]
        ret = segmented_discounted_returns(ep["reward"], ep["seg_last"], ep["seg_id"], boot, self.gamma)
        out = self.update(ep["view"], ep["feature"], ep["action"], ret, ep["prob"])
        if verbose:
            print('[*] PG_LOSS:', np.round(out[0], 6), '/ VF_LOSS:', np.round(out[1], 6), '/ ENT_LOSS:',
                  np.round(out[2], 6), '/ VALUE:', out[3])
        return out

    @torch.no_grad()
    def act(self, **kwargs):
        """multinomial(log policy) (ac.py:43-48, 70): numpy int32 for numpy inputs, a device tensor otherwise."""
        view, feature = kwargs['state'][0], kwargs['state'][1]
        if isinstance(view, torch.Tensor) and view.dtype == torch.bfloat16:       # the engine's bf16 rows [N, 13, 13, 8]
            if getattr(self, "_rollout16", None) is None or self._rollout16_stale:
                self._rollout16, self._rollout16_stale = self.net.bf16_rollout_copy(self.view_space), False
                self._act_graphs = {}
            return self._act_rows16(view, feature)
        v, f = as_tensor(view, self.device), as_tensor(feature, self.device)
        if self.act_autocast is not None and v.is_cuda:
            with torch.autocast("cuda", dtype=self.act_autocast):
                policy = self.net.policy(v, f).float()
        else:
            policy = self.net.policy(v, f)
        action = torch.multinomial(policy, 1, generator=self.generator).reshape(-1).to(torch.int32)
        return action if isinstance(view, torch.Tensor) else action.cpu().numpy()

    def _act_rows16(self, view, feature):
        """Forward of the bf16 twin + the multinomial draw on the engine's observation buffers.  The dense network is
        ~20 small launches per group and the rollout loop is bound by the host issuing them (profiles/r02/
        play_busy_probe.txt), and the engine hands out the SAME buffers every step, so the sequence is captured into a
        CUDA graph per (view buffer, feature buffer) and replayed; the result lands in a fixed tensor that the caller
        copies from right away.  MFMARL_ACT_GRAPH=0, a CPU tensor or a failed capture fall back to the eager calls."""
        def eager():
            policy = self._rollout16.policy(view, feature.to(torch.bfloat16)).float()
            return torch.multinomial(policy, 1, generator=self.generator).reshape(-1).to(torch.int32)

        if not view.is_cuda or not getattr(self, "_act_graph_ok", os.environ.get("MFMARL_ACT_GRAPH", "1") != "0"):
            return eager()
        key = (view.data_ptr(), feature.data_ptr(), tuple(view.shape), tuple(feature.shape))
        hit = self._act_graphs.get(key)
        if hit is None and len(self._act_graphs) >= 4:
            # the caller does not hand in the same buffers again (a graph per call would cost more than it saves)
            self._act_graph_ok, self._act_graphs = False, {}
            return eager()
        if hit is None:
            try:
                side = torch.cuda.Stream(device=view.device)
                side.wait_stream(torch.cuda.current_stream(view.device))
                with torch.cuda.stream(side):
                    eager(); eager()                                   # warm-up outside the capture
                torch.cuda.current_stream(view.device).wait_stream(side)
                graph = torch.cuda.CUDAGraph()
                graph.register_generator_state(self.generator)
                with torch.cuda.graph(graph):
                    out = eager()
                hit = self._act_graphs[key] = (graph, out)
            except Exception as ex:                                    # capture not possible here: stay eager for good
                self._act_graph_ok = False
                print("[mfmarl] rollout graph capture failed (%s): eager rollout forward" % type(ex).__name__)
                return eager()
        hit[0].replay()
        return hit[1]

    @torch.no_grad()
    def _value(self, view, feature, prob):
        v = as_tensor(np.asarray(view)[None] if not isinstance(view, torch.Tensor) else view[None], self.device)
        f = as_tensor(np.asarray(feature)[None] if not isinstance(feature, torch.Tensor) else feature[None], self.device)
        p = None
        if self.use_mf:
            p = as_tensor(np.asarray(prob)[None] if not isinstance(prob, torch.Tensor) else prob[None], self.device)
        return float(self.net(v, f, p)[1][0])

    def losses(self, view, feature, action, ret, prob=None):
        """-> (pg_loss, vf_loss, neg_entropy, value) as tensors (ac.py:82-95)."""
        policy, value = self.net(view, feature, prob)
        advantage = (ret - value).detach()
        log_policy = torch.log(policy + 1e-6)
        log_prob = log_policy.gather(1, action.to(torch.int64).unsqueeze(1)).squeeze(1)
        pg_loss = -(advantage * log_prob).mean()
        vf_loss = self.value_coef * ((ret - value) ** 2).mean()
        neg_entropy = self.ent_coef * (policy * log_policy).sum(dim=1).mean()
        return pg_loss, vf_loss, neg_entropy, value

    def update(self, view, feature, action, ret, prob=None):
        pg_loss, vf_loss, neg_entropy, value = self.losses(view, feature, action, ret, prob)
        self.optimizer.zero_grad(set_to_none=True)
        (pg_loss + vf_loss + neg_entropy).backward()
        if self.grad_sync:
            sync_gradients(self.net.parameters())
        self.optimizer.step()
        self._rollout16_stale = True
        return (float(pg_loss.detach()), float(vf_loss.detach()), float(neg_entropy.detach()),
                float(value.detach().mean()))

    def train(self, verbose=True):
        if self.device_replay is not None and self.device_replay.has_staged:
            return self._train_batched(verbose)
        episodes = list(self.replay_buffer.episodes())
        self.replay_buffer = tools.EpisodesBuffer(use_mean=self.use_mf)
        n = sum(len(ep.rewards) for ep in episodes)
        if n == 0:
            return None
        view = np.empty((n,) + self.view_space, np.float32)
        feature = np.empty((n,) + self.feature_space, np.float32)
        action = np.empty(n, np.int32)
        reward = np.empty(n, np.float32)
        prob = np.zeros((n, self.num_actions), np.float32) if self.use_mf else None
        ct = 0
        for ep in episodes:
            m = len(ep.rewards)
            if self.use_mf:
                assert len(ep.probs) > 0
            keep = self._value(ep.views[-1], ep.features[-1], ep.probs[-1] if self.use_mf else None)
            view[ct:ct + m], feature[ct:ct + m], action[ct:ct + m] = ep.views, ep.features, ep.actions
            reward[ct:ct + m] = discounted_returns(np.array(ep.rewards), keep, self.gamma)
            if self.use_mf:
                prob[ct:ct + m] = ep.probs
            ct += m
        assert n == ct
        out = self.update(as_tensor(view, self.device), as_tensor(feature, self.device),
                          as_tensor(action, self.device, torch.int64), as_tensor(reward, self.device),
                          as_tensor(prob, self.device) if self.use_mf else None)
        if verbose:
            print('[*] PG_LOSS:', np.round(out[0], 6), '/ VF_LOSS:', np.round(out[1], 6), '/ ENT_LOSS:',
                  np.round(out[2], 6), '/ VALUE:', out[3])
        return out

    def save(self, dir_path, step=0):
        import os
        os.makedirs(dir_path, exist_ok=True)
        path = checkpoint_path(dir_path, self.stem, step)
        torch.save({"net": self.net.state_dict(), "optimizer": self.optimizer.state_dict()}, path)
        print("[*] Model saved at: {}".format(path))

    def load(self, dir_path, step=0):
        path = checkpoint_path(dir_path, self.stem, step)
        blob = torch.load(path, map_location=self.device)
        self.net.load_state_dict(blob["net"])
        self.optimizer.load_state_dict(blob["optimizer"])
        print("[*] Loaded model from {}".format(path))
        self._rollout16_stale = True


class ActorCritic(_ActorCriticBase):
    use_mf = False
    stem = "ac"


class MFAC(_ActorCriticBase):
    use_mf = True
    stem = "mfac"

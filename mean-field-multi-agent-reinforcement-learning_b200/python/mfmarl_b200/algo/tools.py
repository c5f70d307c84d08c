"""Replay buffers, scalar logging and the self-play runner of the reference (examples/battle_model/algo/tools.py),
re-expressed without TensorFlow.

Two families live here:

* host buffers with the reference's exact semantics and numpy RNG call order -- `MetaBuffer` (tools.py:26-82),
  `EpisodesBuffer` (:88-173), `AgentMemory` (:178-218), `MemoryGroup` (:223-362) -- used when the reference's
  play loop drives the single-environment binding (`magent.GridWorld`);
* `DeviceMemoryGroup` / `DeviceEpisodesBuffer` (replay_device.py) hold the same data in HBM for the batched engine.

`SummaryObj` (:367-478) keeps its register/write interface but appends JSON lines instead of TF summaries;
`Runner` (:482-651) keeps its arguments minus the TF session, the self-play soft copy is done on the torch
parameters (`tau * right + (1 - tau) * left`, tools.py:566-569).
"""
import json
import os

import numpy as np


class Color:
    INFO = '\033[1;34m{}\033[0m'
    WARNING = '\033[1;33m{}\033[0m'
    ERROR = '\033[1;31m{}\033[0m'


class Buffer:
    def push(self, **kwargs):
        raise NotImplementedError


class MetaBuffer(object):
    """Fixed-capacity ring of rows of one field (tools.py:26-82).  `append` writes at the cursor and wraps;
    `pull` returns the first `length` rows in storage order (not rotated), as the reference does."""

    def __init__(self, shape, max_len, dtype='float32'):
        self.max_len = max_len
        self.data = np.zeros((max_len,) + tuple(shape), dtype=dtype)
        self.length = 0
        self._flag = 0          # write cursor

    def __len__(self):
        return self.length

    def __getitem__(self, idx):
        if not 0 <= idx < self.length:
            raise KeyError()
        return self.data[idx]

    def sample(self, idx):
        return self.data[idx % self.length]

    def pull(self):
        return self.data[:self.length]

    def append(self, value):
        n = len(value)
        room = self.max_len - self._flag
        if n > room:            # fill the tail, continue at the head (a single wrap, like tools.py:66-76)
            self.data[self._flag:] = value[:room]
            rest = n - room
            self.data[:rest] = value[room:]
            self._flag = rest
        else:
            self.data[self._flag:self._flag + n] = value
            self._flag += n
        self.length = min(self.length + n, self.max_len)

    def reset_new(self, start, value):
        self.data[start:] = value


class EpisodesBufferEntry:
    """One agent's trajectory (tools.py:88-117)."""

    def __init__(self):
        self.views, self.features, self.actions, self.rewards, self.probs = [], [], [], [], []
        self.terminal = False

    def append(self, view, feature, action, reward, alive, probs=None):
        self.views.append(view.copy())
        self.features.append(feature.copy())
        self.actions.append(action)
        self.rewards.append(reward)
        if probs is not None:
            self.probs.append(probs)
        if not alive:
            self.terminal = True


class EpisodesBuffer(Buffer):
    """Whole-episode store, one entry per agent id (tools.py:122-173).  `push` walks the agents in a fresh
    np.random.permutation, which fixes the dict (and hence training batch) order."""

    def __init__(self, use_mean=False):
        self.buffer = {}
        self.use_mean = use_mean

    def push(self, **kwargs):
        view, feature = kwargs['state']
        acts, rewards, alives, ids = kwargs['acts'], kwargs['rewards'], kwargs['alives'], kwargs['ids']
        probs = kwargs['prob'] if self.use_mean else None
        order = np.random.permutation(len(view))
        for k in range(len(ids)):
            i = order[k]
            entry = self.buffer.get(ids[i])
            if entry is None:
                entry = self.buffer[ids[i]] = EpisodesBufferEntry()
            entry.append(view[i], feature[i], acts[i], rewards[i], alives[i],
                         probs=probs[i] if self.use_mean else None)

    def reset(self):
        self.buffer = {}

    def episodes(self):
        return self.buffer.values()


class AgentMemory(object):
    """Per-agent rings of length sub_len (tools.py:178-218)."""

    def __init__(self, obs_shape, feat_shape, act_n, max_len, use_mean=False):
        self.obs0 = MetaBuffer(obs_shape, max_len)
        self.feat0 = MetaBuffer(feat_shape, max_len)
        self.actions = MetaBuffer((), max_len, dtype='int32')
        self.rewards = MetaBuffer((), max_len)
        self.terminals = MetaBuffer((), max_len, dtype='bool')
        self.use_mean = use_mean
        if use_mean:
            self.prob = MetaBuffer((act_n,), max_len)

    def append(self, obs0, feat0, act, reward, alive, prob=None):
        self.obs0.append(np.asarray(obs0)[None])
        self.feat0.append(np.asarray(feat0)[None])
        self.actions.append(np.array([act], dtype=np.int32))
        self.rewards.append(np.array([reward]))
        self.terminals.append(np.array([not alive], dtype=bool))
        if self.use_mean:
            self.prob.append(np.asarray(prob)[None])

    def pull(self):
        return {'obs0': self.obs0.pull(), 'feat0': self.feat0.pull(), 'act': self.actions.pull(),
                'rewards': self.rewards.pull(), 'terminals': self.terminals.pull(),
                'prob': self.prob.pull() if self.use_mean else None}


class MemoryGroup(object):
    """Replay ring fed by per-agent trajectories (tools.py:223-362).

    push   -> per-agent memories keyed by id (:281-301)
    tight  -> np.random.shuffle of the ids, then every agent's trajectory is appended to the ring; mask =
              not terminal, and the last row of each trajectory is masked out because its successor in the
              ring belongs to another agent (:262-275, :304-329)
    sample -> np.random.choice(nb_entries, batch); the "next" row is (idx + 1) % nb_entries (:332-350)
    """

    def __init__(self, obs_shape, feat_shape, act_n, max_len, batch_size, sub_len, use_mean=False):
        self.agent = dict()
        self.max_len, self.batch_size, self.sub_len = max_len, batch_size, sub_len
        self.obs_shape, self.feat_shape, self.act_n, self.use_mean = tuple(obs_shape), tuple(feat_shape), act_n, use_mean
        self.obs0 = MetaBuffer(self.obs_shape, max_len)
        self.feat0 = MetaBuffer(self.feat_shape, max_len)
        self.actions = MetaBuffer((), max_len, dtype='int32')
        self.rewards = MetaBuffer((), max_len)
        self.terminals = MetaBuffer((), max_len, dtype='bool')
        self.masks = MetaBuffer((), max_len, dtype='bool')
        if use_mean:
            self.prob = MetaBuffer((act_n,), max_len)
        self._new_add = 0

    def _flush(self, **kwargs):
        self.obs0.append(kwargs['obs0'])
        self.feat0.append(kwargs['feat0'])
        self.actions.append(kwargs['act'])
        self.rewards.append(kwargs['rewards'])
        self.terminals.append(kwargs['terminals'])
        if self.use_mean:
            self.prob.append(kwargs['prob'])
        mask = ~np.asarray(kwargs['terminals'], dtype=bool)
        mask[-1] = False
        self.masks.append(mask)

    def push(self, **kwargs):
        view, feature = kwargs['state']
        for i, _id in enumerate(kwargs['ids']):
            mem = self.agent.get(_id)
            if mem is None:
                mem = self.agent[_id] = AgentMemory(self.obs_shape, self.feat_shape, self.act_n, self.sub_len,
                                                    use_mean=self.use_mean)
            mem.append(obs0=view[i], feat0=feature[i], act=kwargs['acts'][i], reward=kwargs['rewards'][i],
                       alive=kwargs['alives'][i], prob=kwargs['prob'][i] if self.use_mean else None)

    def tight(self):
        ids = list(self.agent.keys())
        np.random.shuffle(ids)
        for _id in ids:
            rows = self.agent[_id].pull()
            self._new_add += len(rows['obs0'])
            self._flush(**rows)
        self.agent = dict()

    def sample(self):
        idx = np.random.choice(self.nb_entries, size=self.batch_size)
        nxt = (idx + 1) % self.nb_entries
        obs, obs_next = self.obs0.sample(idx), self.obs0.sample(nxt)
        feature, feature_next = self.feat0.sample(idx), self.feat0.sample(nxt)
        actions, rewards = self.actions.sample(idx), self.rewards.sample(idx)
        dones, masks = self.terminals.sample(idx), self.masks.sample(idx)
        if self.use_mean:
            return (obs, feature, actions, self.prob.sample(idx), obs_next, feature_next, self.prob.sample(nxt),
                    rewards, dones, masks)
        return obs, feature, obs_next, feature_next, dones, rewards, actions, masks

    def get_batch_num(self, verbose=True):
        if verbose:
            print('\n[INFO] Length of buffer and new add:', len(self.obs0), self._new_add)
        res = self._new_add * 2 // self.batch_size
        self._new_add = 0
        return res

    @property
    def nb_entries(self):
        return len(self.obs0)


class SummaryObj:
    """Scalar log with the reference's interface (tools.py:367-478: register(names), write(dict, step)).
    Instead of TF event files it appends one JSON line per write to <log_dir>/<log_name>/scalars.jsonl."""

    def __init__(self, log_dir, log_name, n_group=1):
        self.name_set = set()
        self.n_group = n_group
        self.dir = os.path.join(log_dir, log_name)
        os.makedirs(self.dir, exist_ok=True)
        self.path = os.path.join(self.dir, "scalars.jsonl")

    def register(self, name_list):
        for name in name_list:
            if name in self.name_set:
                raise Exception("You cannot define different operations with same name: `{}`".format(name))
            self.name_set.add(name)

    def write(self, summary_dict, step):
        assert isinstance(summary_dict, dict)
        row = {"step": int(step)}
        for key, value in summary_dict.items():
            if key not in self.name_set:
                raise Exception("Undefined operation: `{}`".format(key))
            if isinstance(value, list):
                for i in range(self.n_group):
                    row["Agent_%d_%s" % (i, key)] = float(value[i])
            else:
                row["Agent_0_%s" % key] = float(value)
        with open(self.path, "a") as f:
            f.write(json.dumps(row) + "\n")


def soft_copy(dst_vars, src_vars, tau):
    """dst <- (1 - tau) * src + tau * dst  (the self-play op of tools.py:566-569; l_vars = src, r_vars = dst)."""
    import torch
    with torch.no_grad():
        for d, s in zip(dst_vars, src_vars):
            d.mul_(tau).add_(s.to(d.device), alpha=1.0 - tau)


class Runner(object):
    """tools.py:482-651 without the TF session: runs `play_handle` for one round, keeps the kill / reward
    statistics, and in training mode applies the self-play update to the opponent and saves both models
    whenever the main model out-scored it."""

    def __init__(self, env, handles, map_size, max_steps, models, play_handle, render_every=None, save_every=None,
                 tau=None, log_name=None, log_dir=None, model_dir=None, train=False):
        self.env, self.models, self.max_steps, self.handles, self.map_size = env, models, max_steps, handles, map_size
        self.render_every, self.save_every, self.play = render_every, save_every, play_handle
        self.model_dir, self.train, self.tau = model_dir, train, tau
        if self.train:
            self.summary = SummaryObj(log_name=log_name, log_dir=log_dir)
            self.summary_items = ['ave_agent_reward', 'total_reward', 'kill', "Sum_Reward", "Kill_Sum"]
            self.summary.register(self.summary_items)
            assert self.models[0].name != self.models[1].name
            assert len(self.models[0].vars) == len(self.models[1].vars)
            os.makedirs(self.model_dir, exist_ok=True)

    def run(self, variant_eps, iteration, win_cnt=None):
        info = {tag: {'ave_agent_reward': 0., 'total_reward': 0., 'kill': 0.} for tag in ('main', 'opponent')}
        render = (iteration + 1) % self.render_every if self.render_every and self.render_every > 0 else False
        max_nums, nums, agent_r_records, total_rewards = self.play(
            env=self.env, n_round=iteration, map_size=self.map_size, max_steps=self.max_steps, handles=self.handles,
            models=self.models, print_every=50, eps=variant_eps, render=render, train=self.train)
        for i, tag in enumerate(['main', 'opponent']):
            info[tag]['total_reward'] = total_rewards[i]
            info[tag]['kill'] = max_nums[i] - nums[1 - i]
            info[tag]['ave_agent_reward'] = agent_r_records[i]
        if self.train:
            print('\n[INFO] {}'.format(info['main']))
            if info['main']['total_reward'] > info['opponent']['total_reward']:
                print(Color.INFO.format('\n[INFO] Begin self-play Update ...'))
                soft_copy(self.models[1].vars, self.models[0].vars, self.tau)
                print(Color.INFO.format('[INFO] Self-play Updated!\n'))
                print(Color.INFO.format('[INFO] Saving model ...'))
                self.models[0].save(self.model_dir + '-0', iteration)
                self.models[1].save(self.model_dir + '-1', iteration)
                self.summary.write(info['main'], iteration)
        else:
            print('\n[INFO] {0} \n {1}'.format(info['main'], info['opponent']))
            if info['main']['kill'] > info['opponent']['kill']:
                win_cnt['main'] += 1
            elif info['main']['kill'] < info['opponent']['kill']:
                win_cnt['opponent'] += 1
            else:
                win_cnt['main'] += 1
                win_cnt['opponent'] += 1
        return info

"""IL (DQN) and MF-Q on top of ValueNet, with the reference's MemoryGroup replay
(algo/q_learning.py:9-171)."""
from . import base, tools


class DQN(base.ValueNet):
    def __init__(self, name, handle, env, sub_len, memory_size=2 ** 10, batch_size=64, update_every=5, device=None):
        super().__init__(env, handle, name, update_every=update_every, device=device)
        self.replay_buffer = tools.MemoryGroup(self.view_space, self.feature_space, self.num_actions, memory_size,
                                               batch_size, sub_len)

    def flush_buffer(self, **kwargs):
        self.replay_buffer.push(**kwargs)

    def train(self, verbose=True):
        self.replay_buffer.tight()
        losses = []
        for i in range(self.replay_buffer.get_batch_num(verbose)):
            obs, feats, obs_next, feat_next, dones, rewards, actions, masks = self.replay_buffer.sample()
            target_q = self.calc_target_q(obs=obs_next, feature=feat_next, rewards=rewards, dones=dones)
            loss, q = super().train(state=[obs, feats], target_q=target_q, acts=actions, masks=masks)
            self.update()
            losses.append(loss)
            if verbose and i % 50 == 0:
                print("[*] LOSS:", loss, "/ Q:", q)
        return losses

    def save(self, dir_path, step=0):
        self._save(dir_path, "dqn", step)

    def load(self, dir_path, step=0):
        self._load(dir_path, "dqn", step)


class MFQ(base.ValueNet):
    def __init__(self, name, handle, env, sub_len, eps=1.0, update_every=5, memory_size=2 ** 10, batch_size=64,
                 device=None):
        super().__init__(env, handle, name, use_mf=True, update_every=update_every, device=device)
        self.train_ct = 0
        self.replay_buffer = tools.MemoryGroup(self.view_space, self.feature_space, self.num_actions, memory_size,
                                               batch_size, sub_len, use_mean=True)

    def flush_buffer(self, **kwargs):
        self.replay_buffer.push(**kwargs)

    def train(self, verbose=True):
        self.replay_buffer.tight()
        losses = []
        for i in range(self.replay_buffer.get_batch_num(verbose)):
            (obs, feat, acts, act_prob, obs_next, feat_next, act_prob_next, rewards, dones,
             masks) = self.replay_buffer.sample()
            target_q = self.calc_target_q(obs=obs_next, feature=feat_next, rewards=rewards, dones=dones,
                                          prob=act_prob_next)
            loss, q = super().train(state=[obs, feat], target_q=target_q, prob=act_prob, acts=acts, masks=masks)
            self.update()
            losses.append(loss)
            if verbose and i % 50 == 0:
                print("[*] LOSS:", loss, "/ Q:", q)
        return losses

    def save(self, dir_path, step=0):
        self._save(dir_path, "mfq", step)

    def load(self, dir_path, step=0):
        self._load(dir_path, "mfq", step)

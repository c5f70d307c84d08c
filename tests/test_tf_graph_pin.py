"""Pins the PyTorch networks to the reference's TensorFlow graph DEFINITIONS.  TensorFlow 1.x
is not installable here, so the pin is what the graph definition fixes, evaluated two ways that do not go through
torch:

  * a forward pass written in numpy straight from the semantics of the TF layers the reference calls --
    tf.layers.conv2d on NHWC input with an HWIO kernel, VALID padding, stride 1 (base.py:125-136);
    tf.reshape(conv2, [-1, prod]) = row-major over (h, w, c) (:138-139); tf.layers.dense = x @ kernel + bias
    (:141-183); tf.concat axis 1 in the order written; ActorCritic / MFAC the same way (ac.py:50-75, 220-248) --
    fed with TF-LAYOUT weights, against the torch module loaded with the same arrays through
    algo.base.load_tf_variables;
  * hand-computed activations for constant weights, every figure derived in the comments.

Index-dependent weights make any transposed kernel, swapped concat or (c, h, w) flatten show up as a mismatch.
Tolerance 1e-6 relative (north_star) in fp32 against float64 numpy.
"""
import numpy as np
import pytest
import torch

from mfmarl_b200.algo.ac import ACNet, TF_AC_LAYERS, TF_MFAC_LAYERS
from mfmarl_b200.algo.base import QNet, TF_QNET_LAYERS, load_tf_variables

VIEW, FEAT, NACT = (13, 13, 7), (34,), 21


# ------------------------------------------------------------------ TF layer semantics, numpy float64
def tf_conv2d_valid(x, kernel, bias):
    """x [n, h, w, cin], kernel [kh, kw, cin, cout] -> [n, h-kh+1, w-kw+1, cout]"""
    n, h, w, _ = x.shape
    kh, kw, _, cout = kernel.shape
    out = np.zeros((n, h - kh + 1, w - kw + 1, cout))
    for dy in range(kh):
        for dx in range(kw):
            out += np.einsum("nhwi,io->nhwo", x[:, dy:dy + h - kh + 1, dx:dx + w - kw + 1, :], kernel[dy, dx])
    return out + bias


def tf_dense(x, kernel, bias):
    return x @ kernel + bias


relu = lambda a: np.maximum(a, 0.0)


def tf_value_net(v, view, feat, prob, use_mf):
    """ValueNet._construct_net, base.py:123-183"""
    conv2 = relu(tf_conv2d_valid(relu(tf_conv2d_valid(view, v["Conv1/kernel"], v["Conv1/bias"])),
                                 v["Conv2/kernel"], v["Conv2/bias"]))
    flat = conv2.reshape(len(view), -1)                                          # (h, w, c) row-major
    h_obs = relu(tf_dense(flat, v["Dense-Obs/kernel"], v["Dense-Obs/bias"]))
    h_emb = relu(tf_dense(feat, v["Dense-Emb/kernel"], v["Dense-Emb/bias"]))
    cat = np.concatenate([h_obs, h_emb], axis=1)
    if use_mf:
        p = relu(tf_dense(relu(tf_dense(prob, v["Prob-Emb/kernel"], v["Prob-Emb/bias"])),
                          v["Dense-Act-Prob/kernel"], v["Dense-Act-Prob/bias"]))
        cat = np.concatenate([cat, p], axis=1)
    d2 = relu(tf_dense(cat, v["Dense2/kernel"], v["Dense2/bias"]))
    out = relu(tf_dense(d2, v["Dense-Out/kernel"], v["Dense-Out/bias"]))
    return tf_dense(out, v["Q-Value/kernel"], v["Q-Value/bias"])


def tf_actor_critic(v, view, feat, prob, use_mf):
    """ac.py:50-75 (ActorCritic) and :220-248 (MFAC); tf.layers.dense default names dense, dense_1, ..."""
    h_view = relu(tf_dense(view.reshape(len(view), -1), v["dense/kernel"], v["dense/bias"]))
    h_emb = relu(tf_dense(feat, v["dense_1/kernel"], v["dense_1/bias"]))
    cat = np.concatenate([h_view, h_emb], axis=1)
    dense = relu(tf_dense(cat, v["dense_2/kernel"], v["dense_2/bias"]))
    logits = tf_dense(dense / 0.1, v["dense_3/kernel"], v["dense_3/bias"])
    e = np.exp(logits - logits.max(axis=1, keepdims=True))
    policy = np.clip(e / e.sum(axis=1, keepdims=True), 1e-10, 1 - 1e-10)
    if use_mf:
        p = relu(tf_dense(relu(tf_dense(prob, v["dense_4/kernel"], v["dense_4/bias"])), v["dense_5/kernel"], v["dense_5/bias"]))
        value = tf_dense(relu(tf_dense(np.concatenate([cat, p], axis=1), v["dense_6/kernel"], v["dense_6/bias"])),
                         v["dense_7/kernel"], v["dense_7/bias"])
    else:
        value = tf_dense(dense, v["dense_4/kernel"], v["dense_4/bias"])
    return policy, value.reshape(-1)


# ------------------------------------------------------------------ fixtures
def patterned(shape, seed, scale):
    """deterministic, index-dependent, sign-changing weights: w[idx] = scale * sin(seed + 0.37 * flat_index)"""
    n = int(np.prod(shape))
    return (scale * np.sin(seed + 0.37 * np.arange(n))).reshape(shape)


def qnet_variables(use_mf, const=None):
    shapes = {"Conv1": (3, 3, 7, 32), "Conv2": (3, 3, 32, 32), "Dense-Obs": (9 * 9 * 32, 256), "Dense-Emb": (34, 32),
              "Dense2": (288 + (32 if use_mf else 0), 128), "Dense-Out": (128, 64), "Q-Value": (64, NACT)}
    if use_mf:
        shapes.update({"Prob-Emb": (NACT, 64), "Dense-Act-Prob": (64, 32)})
    v = {}
    for k, (name, shape) in enumerate(sorted(shapes.items())):
        fan_in = int(np.prod(shape[:-1]))
        v[name + "/kernel"] = patterned(shape, k, 1.5 / np.sqrt(fan_in)) if const is None else np.full(shape, const[name])
        v[name + "/bias"] = patterned(shape[-1:], 10 + k, 0.1) if const is None else np.zeros(shape[-1:])
    return v


def batch(n, seed):
    rng = np.random.RandomState(seed)
    view = (rng.rand(n, *VIEW) < 0.15) * rng.rand(n, *VIEW)          # sparse like an observation
    return view, rng.rand(n, *FEAT), rng.dirichlet(np.ones(NACT), size=n)


def torch_forward(net, *arrays):
    with torch.no_grad():
        out = net(*[torch.as_tensor(a, dtype=torch.float32) for a in arrays])
    return [o.double().numpy() for o in (out if isinstance(out, tuple) else (out,))]


# ------------------------------------------------------------------ tests
@pytest.mark.parametrize("use_mf", [False, True])
def test_value_net_equals_the_tf_graph_definition(use_mf):
    v = qnet_variables(use_mf)
    net = QNet(VIEW, FEAT, NACT, use_mf)
    load_tf_variables(net, TF_QNET_LAYERS, v)
    view, feat, prob = batch(6, 3)
    want = tf_value_net(v, view, feat, prob, use_mf)
    got, = torch_forward(net, view, feat, prob) if use_mf else torch_forward(net, view, feat)
    assert np.abs(want).max() > 0.05 and want.std() > 1e-3                  # the comparison is not vacuous
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=2e-6 * np.abs(want).max())
    # the orderings matter: a (c, h, w) flatten or a swapped concat gives something else
    wrong = dict(v)
    wrong["Dense-Obs/kernel"] = v["Dense-Obs/kernel"].reshape(9, 9, 32, 256).transpose(2, 0, 1, 3).reshape(2592, 256)
    assert np.abs(tf_value_net(wrong, view, feat, prob, use_mf) - want).max() > 1e-3


def test_value_net_hand_computed_activations():
    """Constant kernels (powers of two, so every figure is exact in fp32), zero biases, all-ones inputs, IL network.  By hand:
    Conv1: 3*3*7 inputs * 1/64 = 63/64 = 0.984375;            Conv2: 3*3*32 * 0.984375 / 64 = 4.4296875   (9 x 9 x 32 of them);
    Dense-Obs: 2592 * 4.4296875 / 4096 = 2.80316162109375;    Dense-Emb: 34 / 64 = 0.53125;
    Dense2: (256 * 2.80316162109375 + 32 * 0.53125) / 1024 = 734.609375 / 1024 = 0.7173919677734375;
    Dense-Out: 128 * 0.7173919677734375 / 64 = 1.434783935546875;   Q-Value: 64 * 1.434783935546875 / 64 = 1.434783935546875."""
    const = {"Conv1": 1 / 64, "Conv2": 1 / 64, "Dense-Obs": 1 / 4096, "Dense-Emb": 1 / 64, "Dense2": 1 / 1024,
             "Dense-Out": 1 / 64, "Q-Value": 1 / 64}
    v = qnet_variables(False, const)
    net = QNet(VIEW, FEAT, NACT, False)
    load_tf_variables(net, TF_QNET_LAYERS, v)
    ones_view, ones_feat = np.ones((2,) + VIEW), np.ones((2,) + FEAT)
    with torch.no_grad():
        c1 = torch.relu(net.conv1(torch.ones(2, 7, 13, 13)))
        c2 = torch.relu(net.conv2(c1))
    assert c1.shape == (2, 32, 11, 11) and c2.shape == (2, 32, 9, 9)         # VALID padding, 3 x 3
    np.testing.assert_allclose(c1.numpy(), 0.984375, rtol=1e-6)
    np.testing.assert_allclose(c2.numpy(), 4.4296875, rtol=1e-6)
    q, = torch_forward(net, ones_view, ones_feat)
    np.testing.assert_allclose(q, 1.434783935546875, rtol=1e-6)
    np.testing.assert_allclose(tf_value_net(v, ones_view, ones_feat, None, False), 1.434783935546875, rtol=1e-12)


@pytest.mark.parametrize("use_mf", [False, True])
def test_actor_critic_equals_the_tf_graph_definition(use_mf):
    shapes = {"dense": (13 * 13 * 7, 256), "dense_1": (34, 256), "dense_2": (512, 512), "dense_3": (512, NACT)}
    if use_mf:
        shapes.update({"dense_4": (NACT, 64), "dense_5": (64, 32), "dense_6": (512 + 32, 256), "dense_7": (256, 1)})
    else:
        shapes["dense_4"] = (512, 1)
    v = {}
    for k, (name, shape) in enumerate(sorted(shapes.items())):
        scale = (0.15 if name == "dense_3" else 1.5) / np.sqrt(shape[0])      # keeps the /0.1 softmax away from saturation
        v[name + "/kernel"], v[name + "/bias"] = patterned(shape, k, scale), patterned(shape[-1:], 20 + k, 0.05)
    net = ACNet(VIEW, FEAT, NACT, use_mf)
    load_tf_variables(net, TF_MFAC_LAYERS if use_mf else TF_AC_LAYERS, v)
    view, feat, prob = batch(5, 8)
    want_policy, want_value = tf_actor_critic(v, view, feat, prob, use_mf)
    got_policy, got_value = torch_forward(net, view, feat, prob) if use_mf else torch_forward(net, view, feat)
    assert want_policy.max() < 0.999 and want_policy.std() > 1e-3 and np.abs(want_value).max() > 0.01
    np.testing.assert_allclose(got_policy, want_policy, rtol=2e-5, atol=1e-7)   # softmax of logits / 0.1 amplifies fp32 rounding
    np.testing.assert_allclose(got_value, want_value, rtol=1e-6, atol=2e-6 * np.abs(want_value).max())


def test_q_target_from_the_tf_graph_definition():
    """base.py:192-220 on top of the TF-semantics forward: target = r + (1 - done) * gamma * Q_target[argmax Q_eval]."""
    from mfmarl_b200.algo import MFQ
    from test_algo_models import FakeEnv
    m = MFQ("mfq", 0, FakeEnv(), 400, device="cpu")
    ve, vt = qnet_variables(True), qnet_variables(True)
    vt = {k: a * 0.9 + 0.01 for k, a in vt.items()}
    load_tf_variables(m.eval_net, TF_QNET_LAYERS, ve)
    load_tf_variables(m.target_net, TF_QNET_LAYERS, vt)
    view, feat, prob = batch(16, 5)
    rng = np.random.RandomState(0)
    rewards, dones = rng.randn(16), rng.rand(16) < 0.3
    e_q, t_q = tf_value_net(ve, view, feat, prob, True), tf_value_net(vt, view, feat, prob, True)
    top2 = np.sort(e_q, axis=1)[:, -2:]
    assert (top2[:, 1] - top2[:, 0]).min() > 1e-4                           # no argmax tie that fp32 could flip
    want = rewards + (1.0 - dones) * t_q[np.arange(16), e_q.argmax(1)] * 0.95
    got = m.calc_target_q(obs=view.astype(np.float32), feature=feat.astype(np.float32), prob=prob.astype(np.float32),
                          rewards=rewards.astype(np.float32), dones=dones).numpy()
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=2e-6 * np.abs(want).max())

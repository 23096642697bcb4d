"""CPU suite, part 5: the DQN learner's host/torch logic on a stand-in env (no GPU)."""
import numpy as np
import torch

from gymwipe_b200 import spaces
from gymwipe_b200.agents.dqn_counter_traffic import CounterTrafficProcessor, DQNLearner, ReplayMemory


class FakeEnv:
    """obs in {-2,0,2}+65536; reward +1 when the agent picks device == (obs > 65536)."""
    COUNTER_BOUND = 65536

    def __init__(self, n):
        self.num_envs = n
        self.device = torch.device("cpu")
        self.action_space = spaces.Dict({"device": spaces.Discrete(2), "duration": spaces.Discrete(20)})
        self.gen = torch.Generator().manual_seed(0)
        self.obs = None

    def reset(self):
        self.obs = 65536 + 2 * torch.randint(-1, 2, (self.num_envs,), generator=self.gen)
        return self.obs

    def step(self, action):
        assert action["device"].dtype == torch.int32 and action["duration"].dtype == torch.int32
        assert int(action["device"].max()) <= 1 and int(action["duration"].max()) <= 19
        reward = (action["device"].long() == (self.obs > 65536).long()).double()
        self.obs = 65536 + 2 * torch.randint(-1, 2, (self.num_envs,), generator=self.gen)
        return self.obs, reward, torch.zeros(self.num_envs, dtype=torch.bool), {}


def test_processor_matches_reference_mapping():
    p = CounterTrafficProcessor(20)
    for a in (0, 19, 20, 39):
        d = p.process_action(a)
        assert d == {"device": a // 20, "duration": a % 20}
    t = p.process_action(torch.tensor([0, 19, 20, 39]))
    assert t["device"].tolist() == [0, 0, 1, 1] and t["duration"].tolist() == [0, 19, 0, 19]


def test_replay_memory_ring():
    m = ReplayMemory(10, "cpu")
    for k in range(4):
        x = torch.arange(4, dtype=torch.float32) + 4 * k
        m.append(x, x.long(), x, x, torch.zeros(4))
    assert m.size == 10 and m.head == 6
    assert sorted(m.obs.tolist()) == [float(v) for v in range(6, 16)]
    o, a, r, n, d = m.sample(5)
    assert o.shape == (5,) and a.dtype == torch.int64


def test_learner_trains_on_fake_env():
    env = FakeEnv(64)
    dqn = DQNLearner(env, nb_steps_warmup=128, normalize_obs=True, lr=1e-2)
    w0 = [p.detach().clone() for p in dqn.model.parameters()]
    hist = dqn.fit(60)
    assert len(hist["loss"]) > 40 and all(l == l for l in hist["loss"])     # finite, training happened
    assert any(not torch.equal(a, b.detach()) for a, b in zip(w0, dqn.model.parameters()))
    assert sum(p.numel() for p in dqn.model.parameters()) == 1 * 16 + 16 + 16 * 16 + 16 + 16 * 16 + 16 + 16 * 40 + 40


def test_acting_and_replay_use_the_same_features():
    """The observation is centred once (``_features``): the network input used to ACT on an observation
    equals the input stored in the replay memory for it (it was centred twice when acting)."""
    env = FakeEnv(8)
    dqn = DQNLearner(env, nb_steps_warmup=10 ** 9, normalize_obs=True)
    seen = []
    model = dqn.model
    dqn.model = lambda x: (seen.append(x.clone()), model(x))[1]
    dqn.fit(3)
    stored = dqn.memory.obs[:24].reshape(3, 8)
    for t in range(3):
        assert torch.equal(seen[t].reshape(-1), stored[t])
    assert set(stored.reshape(-1).tolist()) <= {-2.0, 0.0, 2.0}


def _np_forward(ws, x):
    h = x
    for k in range(4):
        h = h @ ws[2 * k].T + ws[2 * k + 1]
        if k < 3:
            h = np.maximum(h, 0.0)
    return h


def _np_backward(ws, x, a, y):
    """d(mean 0.5 (Q(s,a) - y)^2) / d weights of the 1-16-16-16-40 ReLU MLP, float64."""
    acts, h = [x], x
    pre = []
    for k in range(4):
        z = h @ ws[2 * k].T + ws[2 * k + 1]
        pre.append(z)
        h = np.maximum(z, 0.0) if k < 3 else z
        acts.append(h)
    n = x.shape[0]
    delta = np.zeros_like(acts[-1])
    delta[np.arange(n), a] = (acts[-1][np.arange(n), a] - y) / n
    grads = [None] * 8
    for k in (3, 2, 1, 0):
        grads[2 * k] = delta.T @ acts[k]
        grads[2 * k + 1] = delta.sum(axis=0)
        if k > 0:
            delta = (delta @ ws[2 * k]) * (pre[k - 1] > 0)
    return grads


def test_update_rule_matches_keras_rl_restatement():
    """
    Learner parity: a fixed transition tape through ``DQNLearner.train_on_batch`` against a
    float64 NumPy restatement of the rule keras-rl applies for the reference's agent
    (``agents/dqn_counter_traffic.py:58-70`` -> ``rl/agents/dqn.py`` DQNAgent.backward with the defaults,
    Keras Adam(lr=1e-3), soft target update 1e-2): TD target, loss, weights after three updates, target
    weights, and the Boltzmann action probabilities (``rl/policy.py`` BoltzmannQPolicy: tau 1, clip +-500).
    Tolerance: the learner computes in float32 -- 2e-5 relative on the weights after three Adam steps.
    """
    import numpy as np
    env = FakeEnv(4)
    dqn = DQNLearner(env, normalize_obs=True)
    rs = np.random.RandomState(5)
    for p in dqn.model.parameters():                    # non-degenerate weights (biases are zero at init)
        p.data.add_(torch.as_tensor(rs.uniform(-0.2, 0.2, size=tuple(p.shape)), dtype=torch.float32))
    dqn.target.load_state_dict(dqn.model.state_dict())
    for tp in dqn.target.parameters():
        tp.data.mul_(0.9)
    ws = [p.detach().double().numpy().copy() for p in dqn.model.parameters()]
    wt = [p.detach().double().numpy().copy() for p in dqn.target.parameters()]
    m = [np.zeros_like(w) for w in ws]
    v = [np.zeros_like(w) for w in ws]
    gamma, tau, lr, b1, b2, eps = 0.99, 1e-2, 1e-3, 0.9, 0.999, 1e-7
    for it in range(1, 4):
        obs = rs.choice([-2.0, 0.0, 2.0], size=32)
        nxt = rs.choice([-2.0, 0.0, 2.0], size=32)
        act = rs.randint(0, 40, size=32)
        rew = rs.choice([-2.0, 0.0, 2.0], size=32)
        done = (rs.uniform(size=32) < 0.1).astype(np.float64)
        loss = dqn.train_on_batch(*[torch.as_tensor(a, dtype=torch.float32) for a in (obs, )] +
                                  [torch.as_tensor(act, dtype=torch.int64)] +
                                  [torch.as_tensor(a, dtype=torch.float32) for a in (rew, nxt, done)])
        # --- restatement
        y = rew + gamma * (1.0 - done) * _np_forward(wt, nxt[:, None]).max(axis=1)
        q = _np_forward(ws, obs[:, None])[np.arange(32), act]
        want_loss = np.mean(0.5 * (q - y) ** 2)
        assert abs(float(loss) - want_loss) <= 2e-5 * max(1.0, abs(want_loss))
        g = _np_backward(ws, obs[:, None], act, y)
        lr_t = lr * np.sqrt(1.0 - b2 ** it) / (1.0 - b1 ** it)
        for k in range(8):
            m[k] = b1 * m[k] + (1 - b1) * g[k]
            v[k] = b2 * v[k] + (1 - b2) * g[k] ** 2
            ws[k] = ws[k] - lr_t * m[k] / (np.sqrt(v[k]) + eps)
            wt[k] = tau * ws[k] + (1 - tau) * wt[k]
    for k, (p, tp) in enumerate(zip(dqn.model.parameters(), dqn.target.parameters())):
        assert np.allclose(p.detach().double().numpy(), ws[k], rtol=2e-5, atol=2e-6)
        assert np.allclose(tp.detach().double().numpy(), wt[k], rtol=2e-5, atol=2e-6)
    # Boltzmann policy: p = exp(clip(q / tau, -500, 500)) / sum
    x = np.array([-2.0, 0.0, 2.0])
    qv = _np_forward(ws, x[:, None])
    e = np.exp(np.clip(qv / 1.0, -500.0, 500.0))
    want_p = e / e.sum(axis=1, keepdims=True)
    with torch.no_grad():
        qq = dqn.model(torch.as_tensor(x, dtype=torch.float32).reshape(-1, 1)).double()
        got_p = torch.softmax(torch.clamp(qq / dqn.tau, *dqn.clip), dim=1).numpy()
    assert np.allclose(got_p, want_p, rtol=1e-4, atol=1e-7)
    # and the sampler draws from that distribution
    obs_raw = torch.full((20000,), 65536 + 2, dtype=torch.int64)
    a = dqn.select_action(obs_raw)
    freq = np.bincount(a.numpy(), minlength=40) / 20000.0
    assert np.abs(freq - want_p[2]).max() < 0.02

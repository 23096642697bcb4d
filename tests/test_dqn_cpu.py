"""CPU suite, part 5: the DQN learner's host/torch logic on a stand-in env (no GPU)."""
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

"""
On-device DQN learner for the batched ``CounterTrafficEnv`` -- the caller of the hot path
(mirror of ``agents/dqn_counter_traffic.py`` of the reference, which uses keras-rl on one env).

Same agent design (``agents/dqn_counter_traffic.py:35-77``): a 3x16 ReLU MLP on the scalar
observation with 2*20 = 40 outputs, Boltzmann exploration, Adam(1e-3), soft target update
1e-2, warm-up of 1000 steps, a replay memory of 50 000 transitions per env-batch row, discount
0.99 and batch size 32 (keras-rl's defaults), flat action ``a -> {"device": a // 20,
"duration": a % 20}`` (``CounterTrafficProcessor``, :23-33).  Everything -- replay buffer, policy,
optimiser -- lives on the env's GPU; one ``env.step`` yields ``num_envs`` transitions.

When ``torch.distributed`` is initialised (one process per GPU, each with its own env shard)
gradients are averaged with one flat NCCL all-reduce (~1.4 k parameters) and the per-step
reward statistics come from the step kernel's epilogue through ``StatsReducer``.
"""
import torch
import torch.distributed as dist
from torch import nn

ENV_NAME = 'CounterTraffic-v0'


class CounterTrafficProcessor:
    """``agents/dqn_counter_traffic.py:23-33``: reshapes the flat action into the dict action."""

    def __init__(self, max_duration=20):
        self.max_duration = max_duration

    def process_action(self, flat_action):
        assert flat_action is not None
        if torch.is_tensor(flat_action):
            device = torch.div(flat_action, self.max_duration, rounding_mode="floor")
            duration = flat_action - device * self.max_duration
            return {"device": device.to(torch.int32), "duration": duration.to(torch.int32)}
        device = int(flat_action / self.max_duration)
        duration = flat_action - (device * self.max_duration)
        return {"device": device, "duration": duration}


def build_model(nb_actions, hidden=16):
    """``agents/dqn_counter_traffic.py:46-56``: Dense(16)-ReLU x3, Dense(nb_actions), linear; Keras' Dense
    defaults: glorot-uniform kernels, zero biases."""
    model = nn.Sequential(nn.Linear(1, hidden), nn.ReLU(), nn.Linear(hidden, hidden), nn.ReLU(),
                          nn.Linear(hidden, hidden), nn.ReLU(), nn.Linear(hidden, nb_actions))
    for m in model:
        if isinstance(m, nn.Linear):
            nn.init.xavier_uniform_(m.weight)
            nn.init.zeros_(m.bias)
    return model


class KerasAdam(torch.optim.Optimizer):
    """
    Adam as Keras 2 applies it (``keras/optimizers.py``, the optimiser ``dqn.compile(Adam(lr=1e-3))`` of
    ``agents/dqn_counter_traffic.py:66`` uses): ``lr_t = lr * sqrt(1 - b2^t) / (1 - b1^t)``,
    ``p -= lr_t * m / (sqrt(v) + eps)`` with ``eps = 1e-7`` (``K.epsilon()``) -- the epsilon sits outside the
    bias correction, unlike ``torch.optim.Adam``.
    """

    def __init__(self, params, lr=1e-3, beta_1=0.9, beta_2=0.999, epsilon=1e-7):
        super().__init__(params, dict(lr=lr, beta_1=beta_1, beta_2=beta_2, epsilon=epsilon))
        self.iterations = 0

    @torch.no_grad()
    def step(self):
        self.iterations += 1
        t = self.iterations
        for group in self.param_groups:
            b1, b2 = group["beta_1"], group["beta_2"]
            lr_t = group["lr"] * (1.0 - b2 ** t) ** 0.5 / (1.0 - b1 ** t)
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["m"], st["v"] = torch.zeros_like(p), torch.zeros_like(p)
                st["m"].mul_(b1).add_(p.grad, alpha=1.0 - b1)
                st["v"].mul_(b2).addcmul_(p.grad, p.grad, value=1.0 - b2)
                p.addcdiv_(st["m"], st["v"].sqrt().add_(group["epsilon"]), value=-lr_t)


class ReplayMemory:
    """Ring of (obs, action, reward, next obs, done) tensors on the device."""

    def __init__(self, capacity, device):
        self.capacity = int(capacity)
        self.obs = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.next_obs = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.action = torch.zeros(self.capacity, dtype=torch.int64, device=device)
        self.reward = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.done = torch.zeros(self.capacity, dtype=torch.float32, device=device)
        self.size = 0
        self.head = 0

    def append(self, obs, action, reward, next_obs, done):
        n = obs.numel()
        if n >= self.capacity:
            obs, action, reward, next_obs, done = (x[-self.capacity:] for x in (obs, action, reward, next_obs, done))
            n = self.capacity
        idx = (self.head + torch.arange(n, device=self.obs.device)) % self.capacity
        self.obs[idx], self.action[idx], self.reward[idx] = obs, action, reward
        self.next_obs[idx], self.done[idx] = next_obs, done
        self.head = (self.head + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def sample(self, batch_size, generator=None):
        idx = torch.randint(0, self.size, (batch_size,), device=self.obs.device, generator=generator)
        return self.obs[idx], self.action[idx], self.reward[idx], self.next_obs[idx], self.done[idx]


class DQNLearner:
    """DQN with Boltzmann exploration on a batched env (see module docstring)."""

    def __init__(self, env, hidden=16, lr=1e-3, gamma=0.99, target_model_update=1e-2, nb_steps_warmup=1000,
                 memory_limit=50000, batch_size=32, tau=1.0, clip=(-500.0, 500.0), normalize_obs=False, seed=123):
        self.env = env
        self.device = getattr(env, "device", torch.device("cpu"))
        self.nb_devices = env.action_space.spaces["device"].n
        self.nb_durations = env.action_space.spaces["duration"].n
        self.nb_actions = self.nb_devices * self.nb_durations
        self.processor = CounterTrafficProcessor(self.nb_durations)
        torch.manual_seed(seed)
        self.model = build_model(self.nb_actions, hidden).to(self.device)
        self.target = build_model(self.nb_actions, hidden).to(self.device)
        self.target.load_state_dict(self.model.state_dict())
        self.optimizer = KerasAdam(self.model.parameters(), lr=lr)
        self.gamma, self.tau_update, self.warmup = gamma, target_model_update, nb_steps_warmup
        self.batch_size, self.tau, self.clip = batch_size, tau, clip
        self.obs_center = float(getattr(env, "COUNTER_BOUND", 0)) if normalize_obs else 0.0
        n = getattr(env, "num_envs", 1)
        self.memory = ReplayMemory(max(memory_limit, 4 * n), self.device)
        self.gen = torch.Generator(device=self.device).manual_seed(seed)
        self.step_count = 0
        self.reset_done = True
        self.seed = seed
        self.fused_policy = True        # CUDA envs: action selection through gw_policy_boltzmann (one kernel)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        # warm-up in TRANSITIONS, counted with the smallest shard: shards differ by at most one env
        # (distributed.shard_range), and every rank must start training -- and with it the gradient
        # all-reduce -- in the same iteration
        self.warmup_width = int(n)
        if self.world > 1:
            w = torch.tensor([self.warmup_width], dtype=torch.int64, device=self.device)
            dist.all_reduce(w, op=dist.ReduceOp.MIN)
            self.warmup_width = int(w[0])
        self.history = {"loss": [], "mean_reward": []}

    def _features(self, obs):
        """Network input of a RAW observation: the only place where the observation is centred."""
        return (obs.to(torch.float32) - self.obs_center).reshape(-1, 1)

    def _flat_weights(self):
        """The model's parameters as views into ONE float32 buffer (the layout ``gw_policy_boltzmann`` reads)."""
        if getattr(self, "_flat", None) is None:
            params = list(self.model.parameters())
            flat = torch.cat([p.detach().reshape(-1) for p in params]).contiguous()
            off = 0
            for p in params:
                p.data = flat[off:off + p.numel()].view(p.shape)
                off += p.numel()
            self._flat = flat
        return self._flat

    @torch.no_grad()
    def select_action_fused(self, obs, want_probs=False):
        """
        Action selection in ONE kernel (``gw_policy_boltzmann``): MLP forward, Boltzmann probabilities in float64,
        draw -- instead of ~8 eager launches.  Returns ``(flat int64, {"device", "duration"} int32 action dict[, probs])``;
        the dict is what ``env.step`` takes without further conversion.  CUDA envs with the reference's 3x16 network.
        """
        from gymwipe_b200 import _native as N
        obs = obs.to(torch.int64).contiguous()
        n = obs.numel()
        flat = torch.empty(n, dtype=torch.int64, device=obs.device)
        dev = torch.empty(n, dtype=torch.int32, device=obs.device)
        dur = torch.empty(n, dtype=torch.int32, device=obs.device)
        probs = torch.empty((n, self.nb_actions), dtype=torch.float64, device=obs.device) if want_probs else None
        self._draws = getattr(self, "_draws", 0) + 1
        with torch.cuda.device(obs.device):
            N.check(N.lib().gw_policy_boltzmann(
                self._flat_weights().data_ptr(), self.nb_actions, self.nb_durations, obs.data_ptr(), n,
                float(self.obs_center), float(self.tau), float(self.clip[0]), float(self.clip[1]),
                int(self.seed) & (2 ** 64 - 1), self._draws, int(getattr(getattr(self.env, "_cfg", None), "env_id_offset", getattr(self.env, "env_id_offset", 0))),
                flat.data_ptr(), dev.data_ptr(), dur.data_ptr(), probs.data_ptr() if want_probs else None,
                torch.cuda.current_stream(obs.device).cuda_stream))
        shape = getattr(self.env, "_shape", (n,))
        if n != int(torch.Size(shape).numel()):
            shape = (n,)
        action = {"device": dev.reshape(shape), "duration": dur.reshape(shape)}
        return (flat, action, probs) if want_probs else (flat, action)

    def _fused_ok(self):
        return (self.device.type == "cuda" and 1 <= self.nb_actions <= 160
                and [tuple(p.shape) for p in self.model.parameters()] ==
                [(16, 1), (16,), (16, 16), (16,), (16, 16), (16,), (self.nb_actions, 16), (self.nb_actions,)])

    @torch.no_grad()
    def select_action(self, obs):
        """BoltzmannQPolicy (keras-rl): p ~ exp(clip(q / tau)); ``obs`` is the raw observation."""
        q = self.model(self._features(obs)).double()
        logits = torch.clamp(q / self.tau, self.clip[0], self.clip[1])
        probs = torch.softmax(logits, dim=1)
        return torch.multinomial(probs, 1, generator=self.gen).squeeze(1)

    def _train_step(self):
        return self.train_on_batch(*self.memory.sample(self.batch_size, self.gen))

    def train_on_batch(self, obs, action, reward, next_obs, done):
        """
        One update of keras-rl's ``DQNAgent.backward`` (``rl/agents/dqn.py``; the reference constructs the
        agent with the defaults: no double DQN, no dueling, ``gamma = .99``, ``delta_clip = inf``):
        ``y = r + gamma * (1 - terminal) * max_a Q_target(s', a)``; loss = mean over the batch of
        ``0.5 * (Q(s, a) - y)^2`` (``huber_loss`` with an infinite clip value, masked to the taken action);
        Keras Adam; then the soft target update ``target = tau * model + (1 - tau) * target``
        (``get_soft_target_model_updates``, ``target_model_update = 1e-2``) with the UPDATED weights.
        ``obs`` / ``next_obs`` are network inputs (features), as stored in the replay memory.
        """
        with torch.no_grad():
            target_q = self.target(next_obs.reshape(-1, 1)).max(dim=1).values
            y = reward + self.gamma * (1.0 - done) * target_q
        q = self.model(obs.reshape(-1, 1)).gather(1, action.reshape(-1, 1)).squeeze(1)
        loss = torch.mean(0.5 * (q - y) ** 2)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        if self.world > 1:                              # data-parallel learner: average the gradients
            flat = torch.cat([p.grad.reshape(-1) for p in self.model.parameters()])
            dist.all_reduce(flat, op=dist.ReduceOp.SUM)
            flat /= self.world
            off = 0
            for p in self.model.parameters():
                p.grad.copy_(flat[off:off + p.numel()].reshape(p.shape))
                off += p.numel()
        self.optimizer.step()
        with torch.no_grad():                           # soft target update (target_model_update < 1)
            for tp, p in zip(self.target.parameters(), self.model.parameters()):
                tp.mul_(1.0 - self.tau_update).add_(p, alpha=self.tau_update)
        return loss.detach()

    def fit(self, nb_steps, log_interval=None):
        """Runs ``nb_steps`` batched env steps; returns the history dict."""
        obs = self.env.reset()
        obs = torch.as_tensor(obs, device=self.device).reshape(-1)
        fused = self.fused_policy and self._fused_ok()
        # two result triples used alternately (obs of step t is still needed while step t + 1 writes its own):
        # the loop allocates nothing per step
        bufs = None
        if hasattr(self.env, "_shape_t"):
            bufs = [(torch.empty(self.env._shape_t, dtype=torch.int64, device=self.device),
                     torch.empty(self.env._shape_t, dtype=torch.float64, device=self.device),
                     torch.empty(self.env._shape_t, dtype=torch.bool, device=self.device)) for _ in range(2)]
        for it in range(nb_steps):
            if fused:
                flat, action = self.select_action_fused(obs)
            else:
                flat = self.select_action(obs)
                action = self.processor.process_action(flat)
            if bufs is not None:
                next_obs, reward, done, _ = self.env.step(action, out=bufs[it & 1])
            else:
                next_obs, reward, done, _ = self.env.step(action)
            next_obs = torch.as_tensor(next_obs, device=self.device).reshape(-1)
            reward = torch.as_tensor(reward, device=self.device).reshape(-1)
            done = torch.as_tensor(done, device=self.device).reshape(-1)
            self.memory.append(self._features(obs).squeeze(1), flat, reward.to(torch.float32),
                               self._features(next_obs).squeeze(1), done.to(torch.float32))
            self.step_count += 1
            if self.step_count * self.warmup_width >= self.warmup and self.memory.size >= self.batch_size:
                self.history["loss"].append(self._train_step())
            self.history["mean_reward"].append(reward.double().mean())
            obs = next_obs
            # keras-rl's fit() resets an env whose episode ended (rl/core.py: `if done: ... env.reset()`);
            # CounterTrafficEnv never ends an episode (app. B #1), plant / custom envs may
            if self.reset_done and bool(done.any()):
                ids = done.nonzero().reshape(-1)
                fresh = torch.as_tensor(self.env.reset(env_ids=ids), device=self.device).reshape(-1)
                obs = torch.where(done.bool(), fresh, obs)
            if log_interval and self.step_count % log_interval == 0:
                print("step %d  mean reward %.4f" % (self.step_count, float(self.history["mean_reward"][-1])))
        self.history["loss"] = [float(x) for x in self.history["loss"]]
        self.history["mean_reward"] = [float(x) for x in self.history["mean_reward"]]
        return self.history

    def save_weights(self, path):
        torch.save(self.model.state_dict(), path)

    def load_weights(self, path):
        self.model.load_state_dict(torch.load(path, map_location=self.device))
        self.target.load_state_dict(self.model.state_dict())


def learn(num_envs=4096, nb_steps=500, device="cuda"):
    """``agents/dqn_counter_traffic.py:35-77`` on the batched env."""
    import gymwipe_b200
    env = gymwipe_b200.make(ENV_NAME, num_envs=num_envs, device=device, strict=False)
    env.seed(123)
    dqn = DQNLearner(env)
    dqn.fit(nb_steps, log_interval=max(1, nb_steps // 10))
    return dqn


if __name__ == "__main__":
    learn()

"""GPU suite: the DQN learner drives the batched CUDA env (the caller of the hot path)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_dqn_learner_on_batched_env():
    import gymwipe_b200
    from gymwipe_b200.agents import DQNLearner
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=512, strict=False)
    env.seed(123)
    dqn = DQNLearner(env, nb_steps_warmup=1000)
    hist = dqn.fit(40)
    env.check()                                      # every sampled action was inside the action space
    assert len(hist["mean_reward"]) == 40
    assert len(hist["loss"]) >= 30 and all(l == l and l < float("inf") for l in hist["loss"])
    assert dqn.memory.size == 40 * 512
    assert float(env.stats().cpu()[4]) == 40 * 512   # the step kernel's epilogue counted every env-step
    # raw observations are 65536 + {-2, 0, 2}, fed as a scalar like the reference's input_shape=(1,)
    assert set(dqn.memory.obs[:dqn.memory.size].unique().tolist()) <= {65534.0, 65536.0, 65538.0}


def test_fused_policy_kernel_matches_torch_reference():
    """gw_policy_boltzmann (MLP forward + Boltzmann probabilities + draw in one kernel) against the eager PyTorch
    path: float32 forward, float64 softmax of clip(q / tau) -- tolerance 1e-5 relative on the probabilities (the
    kernel accumulates in a different order than cuBLAS); the draws follow those probabilities; the
    device / duration outputs are the reference's flat-action split (agents/dqn_counter_traffic.py:25-33)."""
    import gymwipe_b200
    from gymwipe_b200.agents import DQNLearner
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=4096, strict=False)
    dqn = DQNLearner(env, normalize_obs=True, tau=0.5)
    g = torch.Generator(device="cuda").manual_seed(3)
    for p in dqn.model.parameters():
        p.data.add_(0.3 * torch.randn(p.shape, generator=g, device="cuda"))
    assert dqn._fused_ok()
    obs = 65536 + 2 * torch.randint(-1, 2, (4096,), generator=g, device="cuda")
    flat, action, probs = dqn.select_action_fused(obs, want_probs=True)
    with torch.no_grad():
        q = dqn.model(dqn._features(obs)).double()
        want = torch.softmax(torch.clamp(q / dqn.tau, dqn.clip[0], dqn.clip[1]), dim=1)
    assert torch.allclose(probs, want, rtol=1e-5, atol=1e-9)
    assert torch.equal(action["device"].long(), flat // 20) and torch.equal(action["duration"].long(), flat % 20)
    assert int(flat.min()) >= 0 and int(flat.max()) < 40
    # weights updated in place by the optimiser are seen by the kernel (parameters are views of the flat buffer)
    for p in dqn.model.parameters():
        p.data.mul_(0.5)
    _, _, probs2 = dqn.select_action_fused(obs, want_probs=True)
    with torch.no_grad():
        want2 = torch.softmax(torch.clamp(dqn.model(dqn._features(obs)).double() / dqn.tau, -500.0, 500.0), dim=1)
    assert torch.allclose(probs2, want2, rtol=1e-5, atol=1e-9)
    # the draws follow the distribution: 200k envs with the same observation
    same = torch.full((200000,), 65538, dtype=torch.int64, device="cuda")
    f, _, pr = dqn.select_action_fused(same, want_probs=True)
    freq = torch.bincount(f, minlength=40).double() / f.numel()
    assert float((freq - pr[0]).abs().max()) < 5e-3
    # a new draw counter gives new samples, the same (seed, counter, env) the same sample
    f2, _ = dqn.select_action_fused(same)
    assert not torch.equal(f, f2)
    dqn._draws -= 1
    f3, _ = dqn.select_action_fused(same)
    assert torch.equal(f2, f3)


def test_learner_fit_uses_the_fused_policy():
    import gymwipe_b200
    from gymwipe_b200.agents import DQNLearner
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=1024, strict=False)
    dqn = DQNLearner(env, nb_steps_warmup=2048, normalize_obs=True)
    hist = dqn.fit(12)
    env.check()
    assert dqn._draws == 12 and len(hist["loss"]) >= 9
    assert set(dqn.memory.obs[:dqn.memory.size].unique().tolist()) <= {-2.0, 0.0, 2.0}


def test_learner_on_a_band_of_three_senders():
    """The learner consumes the general band engine as well: 3 senders x 20 durations = 60 actions through the
    run-time-A policy kernel (probabilities against the eager PyTorch path), a short fit on 512 envs."""
    import numpy as np
    import gymwipe_b200
    from gymwipe_b200.agents import DQNLearner
    from gymwipe_b200.envs import GeneralBandEnv
    sc = {"assignment_duration_factor": 1000, "bands": [{"frequency": 2.4e9, "bandwidth": 22e6, "devices": [
        {"role": "sender", "x": 2.0, "y": 0.0, "mult": 1, "payload": "counter", "interval": 0.001, "dest": 1},
        {"role": "sender", "x": -1.0, "y": 1.7, "mult": 3, "payload": "counter", "interval": 0.001, "dest": 2},
        {"role": "sender", "x": -1.0, "y": -1.7, "mult": 2, "payload": "counter", "interval": 0.001, "dest": 0},
        {"role": "rrm", "x": 0.0, "y": 0.0}]}]}
    env = gymwipe_b200.make('CounterTraffic-v0', num_envs=512, scenario=sc, strict=False)
    assert isinstance(env, GeneralBandEnv)
    dqn = DQNLearner(env, nb_steps_warmup=1000, normalize_obs=True, tau=0.7)
    assert dqn.nb_actions == 60 and dqn._fused_ok()
    g = torch.Generator(device="cuda").manual_seed(5)
    for p in dqn.model.parameters():
        p.data.add_(0.3 * torch.randn(p.shape, generator=g, device="cuda"))
    obs = 65536 + 2 * torch.randint(-1, 2, (2048,), generator=g, device="cuda")
    flat, action, probs = dqn.select_action_fused(obs, want_probs=True)
    with torch.no_grad():
        want = torch.softmax(torch.clamp(dqn.model(dqn._features(obs)).double() / dqn.tau, dqn.clip[0], dqn.clip[1]), dim=1)
    assert torch.allclose(probs, want, rtol=1e-5, atol=1e-9)
    assert torch.equal(action["device"].long(), flat // 20) and torch.equal(action["duration"].long(), flat % 20)
    assert int(flat.min()) >= 0 and int(flat.max()) < 60 and int(action["device"].max()) == 2
    hist = dqn.fit(24)
    env.check()
    assert len(hist["mean_reward"]) == 24 and all(l == l for l in hist["loss"])
    assert int(env.transmissions().sum()) > 24 * 512 and int(env.delivered().sum()) > 0

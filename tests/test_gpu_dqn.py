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

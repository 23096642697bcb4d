"""Learners that consume the batched env (mirror of the reference's ``agents/`` directory)."""
from gymwipe_b200.agents.dqn_counter_traffic import CounterTrafficProcessor, DQNLearner, learn

__all__ = ["CounterTrafficProcessor", "DQNLearner", "learn"]

"""Reward-function descriptors (reference core.py:512-586, reward_functions/weighted_sum_of_errors.py); the reward is
computed in the step kernel's epilogue."""
import warnings

import numpy as np

from .utils import set_state_array


class RewardFunction:
    """reference core.py:512-586"""

    reward_range = (-np.inf, np.inf)

    def set_modules(self, physical_system, reference_generator, constraint_monitor):
        pass

    def reset(self, initial_state=None, initial_reference=None):
        pass

    def close(self):
        pass


class WeightedSumOfErrors(RewardFunction):
    """reference weighted_sum_of_errors.py:9-131:
    r = (1 - d) * (-sum_i w_i (|s_i - s*_i| / l_i)^n_i + b) + d * r_violation"""

    def __init__(self, reward_weights=None, normed_reward_weights=False, violation_reward=None, gamma=0.9, reward_power=1, bias=0.0):
        self._n = reward_power
        self._reward_weights = reward_weights
        self._state_length = None
        self._normed = normed_reward_weights
        self._gamma = gamma
        self._bias = bias
        self._violation_reward = violation_reward

    def set_modules(self, physical_system, reference_generator, constraint_monitor):  # :88-123
        ps = physical_system
        self._state_length = ps.state_space.high - ps.state_space.low
        self._n = set_state_array(self._n, ps.state_names)
        referenced_states = reference_generator.referenced_states
        if self._reward_weights is None:
            names = np.array(ps.state_names)
            sel = names[referenced_states] if np.any(referenced_states) else names
            reward_weights = dict.fromkeys(sel, 1 / len(sel))
        else:
            reward_weights = self._reward_weights
        self._reward_weights = set_state_array(reward_weights, ps.state_names)
        rw_sum = sum(self._reward_weights)
        if rw_sum == 0:
            warnings.warn("All reward weights sum up to zero", Warning, stacklevel=2)
        if self._normed:
            if self._bias == "positive":
                self._bias = 1
            self._reward_weights = self._reward_weights / rw_sum
            self.reward_range = (-1 + self._bias, self._bias)
        else:
            if self._bias == "positive":
                self._bias = rw_sum
            self.reward_range = (-rw_sum + self._bias, self._bias)
        if self._violation_reward is None:
            self._violation_reward = min(self.reward_range[0] / (1.0 - self._gamma), 0)

    def fill_config(self, cfg):
        for i in range(len(self._reward_weights)):
            cfg.reward_weight[i] = float(self._reward_weights[i])
            cfg.reward_power[i] = float(self._n[i])
            cfg.state_length[i] = float(self._state_length[i])
        cfg.reward_bias = float(self._bias)
        cfg.violation_reward = float(self._violation_reward)

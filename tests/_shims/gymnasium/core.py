"""TEST-ONLY minimal `gymnasium.core` (see package docstring)."""


class Env:
    metadata = {}
    render_mode = None
    reward_range = (-float("inf"), float("inf"))
    spec = None
    action_space = None
    observation_space = None

    @property
    def unwrapped(self):
        return self

    def reset(self, *, seed=None, options=None):
        raise NotImplementedError

    def step(self, action):
        raise NotImplementedError

    def render(self):
        pass

    def close(self):
        pass

"""TEST-ONLY minimal `gymnasium.envs.registration` (see package docstring)."""
registry = {}


def register(id, entry_point=None, **kwargs):  # noqa: A002 - mirrors the gymnasium signature
    registry[id] = dict(entry_point=entry_point, kwargs=kwargs.get("kwargs", {}) or {})

"""TEST-ONLY tolerant no-op stand-in for matplotlib (not installed in this image).

Only imported so that the unmodified reference's `import matplotlib...` lines succeed while
`tests/golden/make_golden.py` records golden trajectories (every env there is built with an explicit no-op
visualization) and while tests/agent_surface/examples_harness.py lets the reference's example scripts construct their
MotorDashboard objects; nothing is ever drawn.
"""


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, name):
        return _Anything()

    def __iter__(self):
        return iter(())

    def __getitem__(self, key):  # plt.rcParams["axes.prop_cycle"] and friends
        return _Anything()

    def __len__(self):
        return 0


def use(*a, **k):
    pass


def __getattr__(name):
    return _Anything()

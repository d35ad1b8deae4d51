"""Build the reference's example controllers (examples/classic_controllers, an agent written against gym-electric-motor) for every
registered id and let each compute three actions from fixed synthetic observations — once against the unmodified reference
(`--impl reference`) and once with `gym_electric_motor` aliased to this repo's host package (`--impl b200`).  Prints one JSON object
{env_id: {"ok": bool, "error": str | None, "actions": [[...], ...]}}.  No environment is stepped (no GPU needed): what is compared
is everything the agent READS from the env surface — class hierarchy, spaces, state names, limits, nominal values, motor / load
parameters, tau, index attributes, reference generators.  Container-only: needs /root/reference."""
import importlib
import json
import sys
import types
import warnings

import numpy as np

REF = "/root/reference"
HERE = __file__.rsplit("/", 2)[0]


def main(impl):
    warnings.filterwarnings("ignore")
    sys.dont_write_bytecode = True
    np.complex = complex  # the example predates numpy 2
    sys.path.insert(0, HERE + "/_shims")  # gymnasium / matplotlib stand-ins (absent from this image)
    if impl == "reference":
        sys.path.insert(0, REF + "/src")
        import gym_electric_motor as gem
        from gym_electric_motor.core import ElectricMotorVisualization

        class NoViz(ElectricMotorVisualization):
            pass

        import gymnasium

        ids = sorted(k for k in gymnasium.envs.registration.registry.keys() if "-v0" in k and ("Cont-" in k or "Finite-" in k))
        make = lambda env_id: gem.make(env_id, visualization=NoViz())  # noqa: E731
    else:
        sys.path.insert(0, HERE.rsplit("/", 1)[0])
        import gym_electric_motor_b200 as gem

        gem.install_as_gym_electric_motor()
        # plotting helpers of the example (out of scope): inert stand-ins so that its modules import
        for name, classes in (("gym_electric_motor.visualization.motor_dashboard_plots", ("StatePlot", "TimePlot")),
                              ("gym_electric_motor.visualization.motor_dashboard_plots.base_plots", ("TimePlot",)),
                              ("gym_electric_motor.visualization.render_modes", ("RenderMode",))):
            m = types.ModuleType(name)
            for c in classes:
                setattr(m, c, type(c, (), {"__init__": lambda self, *a, **k: None}))
            sys.modules[name] = m
        ids = sorted(gem.env_ids())
        make = gem.make
    sys.path.insert(0, REF + "/examples/classic_controllers")
    from classic_controllers import Controller

    out = {}
    for env_id in ids:
        rec = dict(ok=False, error=None, actions=[])
        try:
            env = make(env_id)
            ctrl = Controller.make(env)
            rec["ok"] = True
            rng = np.random.default_rng(7)
            n_state, n_ref = len(env.state_names), len(env.reference_generator.reference_names)
            for _ in range(3):
                state = rng.uniform(-0.5, 0.5, n_state)
                ref = rng.uniform(-0.3, 0.3, n_ref)
                rec["actions"].append(np.atleast_1d(np.asarray(ctrl.control(state, ref), dtype=float)).ravel().tolist())
        except Exception as e:  # the example itself is stale for some ids (same failure on both sides)
            rec["error"] = f"{type(e).__name__}: {str(e)[:80]}"
        out[env_id] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[sys.argv.index("--impl") + 1] if "--impl" in sys.argv else "b200")

"""Run the reference's example SCRIPTS (examples/environment_features/*.py, examples/classic_controllers/*_example.py) and the
`gem.make` cells of its notebooks (reinforcement_learning_controllers/*.ipynb, model_predictive_controllers/*.ipynb), unmodified, up to
and including their `env = gem.make(...)` call — once against the unmodified reference (`--impl reference`) and once with
`gym_electric_motor` aliased to this repo's host package (`--impl b200`, which then also derives the C-ABI config with
`env.build_config()`).  For every environment that could be built a SUMMARY of what the user's kwargs produced is printed — state and
reference names, limits, nominal state, spaces, tau, reward weights, motor parameters, supply voltage — so the two sides can be
compared field by field.  No stepping, no GPU.  Output: one JSON object {script: {"verdict": "ok" | "<error>", "summary": {...}}}.
Container-only: needs /root/reference."""
import ast
import glob
import json
import os
import sys
import traceback
import types
import warnings

import numpy as np

REF = "/root/reference"
HERE = __file__.rsplit("/", 2)[0]


def summary(env):
    ps = env.physical_system.unwrapped
    sp = env.action_space
    return dict(
        state_names=list(env.state_names), reference_names=list(env.reference_generator.reference_names),
        limits=[float(v) for v in env.limits], nominal_state=[float(v) for v in env.nominal_state], tau=float(ps.tau),
        state_low=[float(v) for v in env.observation_space.spaces[0].low], state_high=[float(v) for v in env.observation_space.spaces[0].high],
        ref_low=[float(v) for v in env.observation_space.spaces[1].low], ref_high=[float(v) for v in env.observation_space.spaces[1].high],
        action=[type(sp).__name__, [int(v) for v in np.atleast_1d(getattr(sp, "nvec", getattr(sp, "n", 0)))] if not hasattr(sp, "low")
                else [[float(v) for v in sp.low], [float(v) for v in sp.high]]],
        motor_parameter={k: float(v) for k, v in sorted(ps.electrical_motor.motor_parameter.items()) if np.ndim(v) == 0},
        j_total=float(ps.mechanical_load.j_total), u_sup=float(ps.supply.u_nominal),
        reward_weights=[float(v) for v in np.asarray(env.reward_function._reward_weights, dtype=float)],
        reward_range=[float(v) for v in env.reward_function.reward_range],
    )


def main(impl):
    warnings.filterwarnings("ignore")
    sys.dont_write_bytecode = True
    np.complex = complex  # the examples predate numpy 2
    sys.path.insert(0, HERE + "/_shims")  # gymnasium / matplotlib stand-ins (absent from this image)
    if impl == "reference":
        sys.path.insert(0, REF + "/src")
        import gym_electric_motor  # noqa: F401
        from gym_electric_motor.visualization import MotorDashboard

        # nothing is drawn in this harness and the matplotlib stand-in cannot carry a real dashboard: hollow the class out in place (the env
        # classes hold a reference to it for their default visualization), constructor kwargs are kept for the agents that read them
        def _init(self, *a, update_interval=1000, **k):
            self._update_interval = update_interval  # read back through the class's own `update_interval` property

        for cls in MotorDashboard.__mro__:
            if cls.__module__.startswith("gym_electric_motor.visualization"):
                for name, attr in list(vars(cls).items()):
                    if callable(attr) and not name.startswith("__"):
                        setattr(cls, name, lambda self, *a, **k: None)
                cls.__init__ = _init
    else:
        sys.path.insert(0, HERE.rsplit("/", 1)[0])
        import gym_electric_motor_b200 as gemb

        gemb.install_as_gym_electric_motor()
        # plotting helpers the example modules import (out of scope here): inert stand-ins so that those modules load
        for name, classes in (("gym_electric_motor.visualization.motor_dashboard_plots", ("StatePlot", "TimePlot", "MeanEpisodeRewardPlot")),
                              ("gym_electric_motor.visualization.motor_dashboard_plots.base_plots", ("TimePlot",)),
                              ("gym_electric_motor.visualization.render_modes", ("RenderMode",))):
            m = types.ModuleType(name)
            for c in classes:
                setattr(m, c, type(c, (), {"__init__": lambda self, *a, **k: None, "__getattr__": lambda self, n: {} if n.endswith("_cfg") else None}))
            sys.modules[name] = m
        sys.modules["gym_electric_motor.visualization.render_modes"].RenderMode = types.SimpleNamespace(Figure="figure", FigureOnce="figure_once")
    sys.path.insert(0, REF + "/examples/classic_controllers")

    def finish(env):
        if impl == "b200":
            env.build_config()
        return summary(env)

    def is_make(node, names=("env",)):
        return isinstance(node, ast.Assign) and "Attribute(value=Name(id='gem', ctx=Load()), attr='make'" in ast.dump(node.value)

    out = {}
    # ---- scripts: module level + the __main__ block up to the gem.make assignment
    for f in sorted(glob.glob(REF + "/examples/environment_features/*.py") + glob.glob(REF + "/examples/classic_controllers/*_example.py")):
        body, made_at_module_level = [], False
        for node in ast.parse(open(f).read()).body:
            if isinstance(node, ast.If) and "__name__" in ast.dump(node.test):
                for sub in node.body:
                    body.append(sub)
                    if is_make(sub):
                        break
            else:
                body.append(node)
                if is_make(node):  # a script that simulates at import time: stop right behind its gem.make
                    made_at_module_level = True
                    break
        ns = {"__name__": "example", "__file__": f}
        rec = dict(verdict="ok", summary=None, make_at_module_level=made_at_module_level)
        try:
            exec(compile(ast.Module(body=body, type_ignores=[]), f, "exec"), ns)
            rec["summary"] = finish(ns["env"])
        except Exception as e:
            tb = traceback.extract_tb(e.__traceback__)[-1]
            rec["verdict"] = f"{type(e).__name__}: {str(e)[:160]} ({tb.filename.split('/')[-1]}:{tb.lineno})"
        out[os.path.basename(f)] = rec
    # ---- notebooks: every code cell up to the one that calls gem.make; statements that need packages absent from this image
    # (stable_baselines3, gekko, gymnasium.wrappers) are skipped one by one, the gem imports / parameter definitions around them run
    for f in sorted(glob.glob(REF + "/examples/reinforcement_learning_controllers/*.ipynb") + glob.glob(REF + "/examples/model_predictive_controllers/*.ipynb")):
        cells = ["".join(c["source"]) for c in json.load(open(f))["cells"] if c["cell_type"] == "code"]
        ns, rec = {}, dict(verdict="no gem.make cell", summary=None)
        for i, src in enumerate(cells):
            src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("%", "!")))
            try:
                nodes = ast.parse(src).body
            except SyntaxError:
                continue
            done = False
            for node in nodes:
                try:
                    exec(compile(ast.Module(body=[node], type_ignores=[]), f"cell{i}", "exec"), ns)
                    if is_make(node):
                        rec["summary"] = finish(ns[node.targets[0].id])
                        rec["verdict"] = "ok"
                except Exception as e:
                    if is_make(node):
                        rec["verdict"] = f"{type(e).__name__}: {str(e)[:160]}"
                if is_make(node):
                    done = True
                    break
            if done:
                break
        out[os.path.basename(f)] = rec
    print(json.dumps(out))


if __name__ == "__main__":
    main(sys.argv[sys.argv.index("--impl") + 1] if "--impl" in sys.argv else "b200")

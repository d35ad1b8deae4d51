"""Run the reference's example SCRIPTS (examples/environment_features/*.py, examples/classic_controllers/*_example.py) unmodified up to and
including their `env = gem.make(...)` call, with `gym_electric_motor` aliased to this repo's host package, then derive the C-ABI config
(`env.build_config()`): every kwarg a user script passes (initialisers, switched reference generators, external speed profiles,
solvers by submodule path, dashboards, `Motor(...).env_id()`) must be accepted.  No stepping, no GPU.  Prints one JSON object
{script: "ok" | "<error>"}.  Container-only: needs /root/reference."""
import ast, importlib, json, sys, types, warnings, traceback, glob, os
warnings.filterwarnings("ignore"); sys.dont_write_bytecode = True
import numpy as np; np.complex = complex
HERE = __file__.rsplit("/", 2)[0]
sys.path.insert(0, HERE + "/_shims"); sys.path.insert(0, HERE.rsplit("/", 1)[0])
import gym_electric_motor_b200 as gemb
gemb.install_as_gym_electric_motor()
for name, classes in (("gym_electric_motor.visualization.motor_dashboard_plots", ("StatePlot", "TimePlot", "MeanEpisodeRewardPlot")),
                      ("gym_electric_motor.visualization.motor_dashboard_plots.base_plots", ("TimePlot",)),
                      ("gym_electric_motor.visualization.render_modes", ("RenderMode",))):
    m = types.ModuleType(name)
    for c in classes:
        setattr(m, c, type(c, (), {"__init__": lambda self, *a, **k: None, "__getattr__": lambda self, n: {} if n.endswith("_cfg") else None}))
    sys.modules[name] = m
sys.modules["gym_electric_motor.visualization.render_modes"].RenderMode = types.SimpleNamespace(Figure="figure", FigureOnce="figure_once")
sys.path.insert(0, "/root/reference/examples/classic_controllers")
files = sorted(glob.glob("/root/reference/examples/environment_features/*.py") + glob.glob("/root/reference/examples/classic_controllers/*_example.py"))
out = {}
for f in files:
    tree = ast.parse(open(f).read())
    body = []
    for node in tree.body:
        if isinstance(node, ast.If) and "__name__" in ast.dump(node.test):
            for sub in node.body:
                body.append(sub)
                if isinstance(sub, ast.Assign) and "make" in ast.dump(sub.value) and any(getattr(t, "id", "") == "env" for t in sub.targets):
                    break
        else:
            body.append(node)
    mod = ast.Module(body=body, type_ignores=[])
    ns = {"__name__": "example", "__file__": f}
    try:
        exec(compile(mod, f, "exec"), ns)
        env = ns["env"]
        cfg = env.build_config()
        out[os.path.basename(f)] = "ok"
    except Exception as e:
        tb = traceback.extract_tb(e.__traceback__)[-1]
        out[os.path.basename(f)] = f"{type(e).__name__}: {str(e)[:160]} ({tb.filename.split('/')[-1]}:{tb.lineno})"


# ---- notebooks: every code cell up to the one that calls gem.make; statements that need packages absent from this image (stable_baselines3,
# gekko, gymnasium.wrappers) are skipped one by one, the gem imports / parameter definitions around them run
NOTEBOOKS = sorted(glob.glob("/root/reference/examples/reinforcement_learning_controllers/*.ipynb") + glob.glob("/root/reference/examples/model_predictive_controllers/*.ipynb"))
for f in NOTEBOOKS:
    cells = ["".join(c["source"]) for c in json.load(open(f))["cells"] if c["cell_type"] == "code"]
    ns = {}
    verdict = "no gem.make cell"
    for i, src in enumerate(cells):
        src = "\n".join(l for l in src.splitlines() if not l.lstrip().startswith(("%", "!")))
        try:
            nodes = ast.parse(src).body
        except SyntaxError:
            continue
        made = None
        for node in nodes:
            is_make = isinstance(node, ast.Assign) and "Attribute(value=Name(id='gem', ctx=Load()), attr='make'" in ast.dump(node.value)
            try:
                exec(compile(ast.Module(body=[node], type_ignores=[]), f"cell{i}", "exec"), ns)
                if is_make:
                    made = ns[node.targets[0].id]
            except Exception as e:
                if is_make:
                    tb = traceback.extract_tb(e.__traceback__)[-1]
                    verdict = f"{type(e).__name__}: {str(e)[:160]} ({tb.filename.split('/')[-1]}:{tb.lineno})"
                    made = False
            if is_make:
                break
        if made is not None:
            if made is not False:
                try:
                    made.build_config()
                    verdict = "ok"
                except Exception as e:
                    verdict = f"{type(e).__name__}: {str(e)[:160]}"
            break
    out[os.path.basename(f)] = verdict
print(json.dumps(out))

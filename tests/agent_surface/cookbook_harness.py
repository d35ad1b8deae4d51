"""Execute the code cells of the reference's examples/environment_features/GEM_cookbook.ipynb (the tutorial that walks through every
`gem.make` kwarg: supply, motor parameters / limits / initialiser, load initialiser, reward function, reference generators, dashboard,
solver, physical-system wrappers, constraints) against this repo's host package aliased as `gym_electric_motor`, up to `gem.make` +
`build_config()`.  User-defined PYTHON components of the tutorial (a Constraint subclass, a function constraint, a custom wrapper,
a file-writing callback) are host code: the harness checks that the constraints are REJECTED loudly and goes on with the built-in
ones.  Prints one JSON object.  Container-only: needs /root/reference."""
import importlib, json, sys, types, warnings
warnings.filterwarnings("ignore"); sys.dont_write_bytecode=True
HERE = __file__.rsplit("/", 2)[0]
sys.path.insert(0, HERE + "/_shims"); sys.path.insert(0, HERE.rsplit("/", 1)[0])
import gym_electric_motor_b200 as gemb
gemb.install_as_gym_electric_motor()
nb = json.load(open('/root/reference/examples/environment_features/GEM_cookbook.ipynb'))
cells = [''.join(c['source']) for c in nb['cells'] if c['cell_type'] == 'code']
ns = {}
def run(i, src=None):
    src = src if src is not None else cells[i]
    src = "\n".join(l for l in src.splitlines() if not l.startswith("%"))
    exec(compile(src, f"cell{i}", "exec"), ns)
run(0)                                   # gem.make + env.visualizations[0].initialize()
for i in (2, 3, 4, 5, 6, 7, 8, 9, 10, 12): run(i)
ns["my_callback"] = []                   # cell 11 writes files: skipped
out = dict(constraints=[type(c).__name__ for c in ns["constraints"]])
try:
    run(15)
    out["custom_constraints"] = "accepted"
except Exception as e:
    out["custom_constraints"] = f"{type(e).__name__}: {str(e)[:120]}"
ns["constraints"] = [c for c in ns["constraints"] if isinstance(c, gemb.constraints.SquaredConstraint) or isinstance(c, str)]
run(15)
cfg = ns["env"].build_config()
import gymnasium  # the stand-in of tests/_shims here; the package subclasses whatever `gymnasium.Env` is importable
out["is_gymnasium_env"] = isinstance(ns["env"], gymnasium.Env) and isinstance(ns["env"].action_space, gymnasium.spaces.Discrete)
out.update(env_class=type(ns["env"]).__name__, n_state_ops=cfg.n_state_ops, n_ref=cfg.n_ref, init_random=cfg.init_random, tau=cfg.tau, u_sup=cfg.u_sup,
           solver_kind=cfg.solver_kind, state_names=list(ns["env"].state_names), reward_i_sq=cfg.reward_weight[ns["env"].state_names.index("i_sq")])
print(json.dumps(out))

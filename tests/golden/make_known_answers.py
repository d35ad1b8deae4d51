#!/usr/bin/env python
"""Dump the KNOWN-ANSWER VECTORS of the reference's own unit tests for the hot path into
tests/golden/known_answers.json, so that the CPU oracle can be pinned to them on any box.

Runs only in the build container: it imports the reference's TEST modules from /root/reference/tests (their tables live in module
globals and pytest.mark.parametrize arguments) and, for the converter tables, replays the reference tests' own call protocol
against the reference implementation first (so the recorded call sequences are known to satisfy the tables).  Nothing is copied
by hand.  Sources:
  tests/test_physical_systems/test_converters.py:14-257 (tables), :260-300, :313-367 (finite 1QC/2QC/4QC protocol),
      :419-503 (continuous 1QC/2QC/4QC, comparable_voltage), :592-640 (finite B6 bridge, per-leg table)
  tests/test_physical_systems/test_mechanical_loads.py:191-211 (PolynomialStaticLoad.mechanical_ode known answers)
  tests/test_physical_systems/test_solvers.py:248-269 (EulerSolver one-step / n-step known answers on tests/conf.py:418-434)
  tests/test_constraints/test_limit_constraint.py:33-66, test_squared_constraint.py:25-99 (truth tables)
  tests/test_reward_functions/test_weighted_sum_of_errors.py:150-218 (reward cases)
"""
import json
import os
import sys
import warnings
from random import seed, uniform

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.dont_write_bytecode = True
try:
    import gymnasium  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(HERE, "..", "_shims"))
sys.path.insert(0, "/root/reference/src")
sys.path.insert(0, "/root/reference")  # the reference's `tests` package
warnings.filterwarnings("ignore")

import tests.test_physical_systems.test_converters as tc  # noqa: E402
import tests.test_physical_systems.test_mechanical_loads as tl  # noqa: E402
import tests.test_constraints.test_limit_constraint as tlc  # noqa: E402
import tests.test_constraints.test_squared_constraint as tsc  # noqa: E402
import tests.test_reward_functions.test_weighted_sum_of_errors as tw  # noqa: E402
import gym_electric_motor.physical_systems.converters as cv  # noqa: E402


def params_of(func, name_hint):
    """argument lists of the pytest.mark.parametrize decorator whose names contain `name_hint`"""
    for m in func.pytestmark:
        names = m.args[0] if isinstance(m.args[0], (list, tuple)) else [s.strip() for s in m.args[0].split(",")]
        if name_hint in names:
            return list(names), list(m.args[1])
    raise KeyError(name_hint)


def finite_qc():
    out = []
    table = [("1QC", cv.FiniteOneQuadrantConverter, tc.g_actions_1qc, tc.g_i_ins_1qc, tc.g_1qc_test_voltages),
             ("2QC", cv.FiniteTwoQuadrantConverter, tc.g_actions_2qc, tc.g_i_ins_2qc, tc.g_2qc_test_voltages),
             ("4QC", cv.FiniteFourQuadrantConverter, tc.g_actions_4qc, tc.g_i_ins_4qc, tc.g_4qc_test_voltages)]
    for kind, cls, actions, i_ins, tv in table:
        for tau in tc.g_taus:
            for il_factor in tc.g_interlocking_times:
                il = float(il_factor * tau)
                conv = cls(tau=tau, interlocking_time=il)
                assert conv.reset() == [0.0]
                times = tc.g_times_4qc * tau  # (sic) every finite converter test uses the 4QC time grid, :333/:349
                calls, k = [], 0
                for t, a, i_in in zip(times, actions, i_ins):
                    steps = conv.set_action(int(a), float(t))
                    for ts in steps:
                        u = conv.convert([float(i_in)], float(ts))
                        expected = float((tv[1] if il > 0 else tv[0])[k])
                        assert u == [expected], (kind, tau, il, k)  # the reference passes its own table
                        calls.append(dict(action=int(a), t_set=float(t), i_in=float(i_in), t_conv=float(ts), expected=expected))
                        k += 1
                out.append(dict(kind=kind, tau=float(tau), interlocking_time=il, calls=calls))
    return out


def finite_b6():
    """:592-640, first part: default-initialised bridge, each leg against u_out"""
    tau = tc.cf.converter_parameter["tau"]
    actions = [[4, 5, 6, 7, 0, 1, 2, 5, 3, 6], [2, 3, 6, 7, 0, 1, 4, 2, 5, 6], [1, 3, 5, 7, 0, 2, 4, 3, 6, 5]]
    i_ins = [0.5, 0, -0.5, 0.5, 0.5, 0, -0.5, -0.5, 0.5, 0.5]
    u_out = [1, 1, 1, 1, -1, -1, -1, 1, -1, 1]
    # the literals above are the test's own (:614-620); replay them against the reference before recording
    conv = cv.FiniteB6BridgeConverter()
    out = []
    for k in range(3):
        conv.reset()
        i_in = [[0.5], [0], [-0.5]]
        calls, step = [], 0
        for t, a, ii in zip(np.arange(10) * tau, actions[k], i_ins):
            for ts in conv.set_action(a, t):
                i_in[k] = [ii]
                u = conv.convert(i_in, ts)
                assert u[k] == 0.5 * u_out[step]
                calls.append(dict(action=int(a), t_set=float(t), i_in=[float(x[0]) for x in i_in], t_conv=float(ts), leg=k, expected=0.5 * u_out[step]))
                step += 1
        out.append(dict(tau=1e-5, interlocking_time=0.0, leg=k, calls=calls))
    return out


def cont_qc():
    out = []
    for kind, cls in (("1QC", cv.ContOneQuadrantConverter), ("2QC", cv.ContTwoQuadrantConverter), ("4QC", cv.ContFourQuadrantConverter)):
        for tau in tc.g_taus:
            for il_factor in tc.g_interlocking_times:
                il = float(il_factor * tau)
                conv = cls(tau=tau, interlocking_time=il)
                assert conv.reset() == [0.0]
                seed(123)  # :436-437
                actions = [[uniform(conv.action_space.low, conv.action_space.high)] for _ in range(len(tc.g_times_cont))]
                calls = []
                for idx, t in enumerate(tc.g_times_cont * tau):
                    a = actions[idx]
                    for ts in conv.set_action(a, t):
                        for i_in in tc.g_i_ins_cont:
                            if kind == "1QC":
                                i_in = abs(i_in)
                            u = conv.convert([i_in], ts)
                            exp = tc.comparable_voltage(cls, a[0], i_in, tau, il, None)
                            assert abs(float(np.asarray(exp).ravel()[0]) - u[0]) < 1e-5
                            calls.append(dict(action=float(np.asarray(a[0]).ravel()[0]), t_set=float(t), i_in=float(i_in), t_conv=float(ts),
                                              expected=float(np.asarray(exp).ravel()[0])))
                out.append(dict(kind=kind, tau=float(tau), interlocking_time=il, calls=calls))
    return out


def poly_load():
    names, cases = params_of(tl.test_PolynomialStaticLoad_MechanicalOde, "omega")
    return dict(load_parameter=dict(j_load=1e-4, a=0.01, b=0.02, c=0.03), torque=2.0,  # literals of the test body :205-207
                cases=[dict(omega=float(c[0]), expected=float(c[1])) for c in cases])


def euler_solver():
    """test_solvers.py:248-269 (TestEulerSolver.test_private_integration): system = tests/conf.py:418-434, y0 = [1, 6], tau = 1e-3, u = 2"""
    import tests.test_physical_systems.test_solvers as ts

    names, cases = params_of(ts.TestEulerSolver.test_private_integration, "expected_state")
    out = []
    for nsteps, expected in cases:
        sol = ts.EulerSolver(nsteps=nsteps)
        sol.set_system_equation(ts.system, ts.jacobian)
        sol.set_initial_value(ts.TestEulerSolver._state, ts.TestEulerSolver._t)
        sol.set_f_params(2)
        got = sol.integrate(ts.TestEulerSolver._t + 1e-3)
        assert sum(abs(got - expected)) < 1e-6
        out.append(dict(nsteps=int(nsteps), y0=[float(v) for v in ts.TestEulerSolver._state], tau=1e-3, u=2.0, expected=[float(v) for v in expected],
                        reference_result=[float(v) for v in got]))
    return out


def constraints():
    out = []
    for mod, cls_name, kind in ((tlc, "TestLimitConstraint", "limit"), (tsc, "TestSquaredConstraint", "squared")):
        cls = getattr(mod, cls_name)
        names, cases = params_of(cls.test_call, "expected_violation")
        for c in cases:
            d = dict(zip(names, c))
            n = len(d["ps"].state_names)
            obs = d["observed_state_names"]
            idx = list(range(n)) if "all_states" in obs else [d["ps"].state_names.index(s) for s in obs]
            out.append(dict(kind=kind, n_state=n, observed=idx, state=[float(v) for v in d["state"]], expected=float(d["expected_violation"])))
    return out


def wse_rewards():
    names, cases = params_of(tw.TestWeightedSumOfErrors.test_reward, "expected_rw")
    out = []
    ps_, rg_, cm_ = tw.DummyPhysicalSystem(state_length=3), tw.DummyReferenceGenerator(), tw.DummyConstraintMonitor()  # the test's own dummies :197-199
    for c in cases:
        d = dict(zip(names, c))
        rg_.set_modules(ps_)
        rf = tw.TestWeightedSumOfErrors.class_to_test(reward_weights=d["reward_weights"], bias=d["bias"], violation_reward=d["violation_reward"])
        rf.set_modules(ps_, rg_, cm_)
        assert rf.reward(d["state"], d["reference"], violation_degree=d["violation_degree"]) == d["expected_rw"]
        out.append(dict(state_length=[float(v) for v in rf._state_length], reward_power=[float(v) for v in np.broadcast_to(rf._n, (3,))],
                        reward_weights=[float(v) for v in d["reward_weights"]], violation_reward=float(d["violation_reward"]), bias=float(d["bias"]),
                        violation_degree=float(d["violation_degree"]), state=[float(v) for v in d["state"]],
                        reference=[float(v) for v in d["reference"]], expected=float(d["expected_rw"])))
    return out


if __name__ == "__main__":
    ka = dict(finite_qc=finite_qc(), finite_b6=finite_b6(), cont_qc=cont_qc(), poly_load=poly_load(), euler=euler_solver(), constraints=constraints(),
              wse_rewards=wse_rewards())
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f)
    print({k: (len(v) if isinstance(v, list) else len(v["cases"])) for k, v in ka.items()},
          "convert calls:", sum(len(c["calls"]) for k in ("finite_qc", "finite_b6", "cont_qc") for c in ka[k]))

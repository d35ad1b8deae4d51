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
  tests/test_physical_system_wrappers/test_dq_to_abc_action_processor.py:27-53 (dq -> abc known answers, advanced angle)
  tests/test_physical_system_wrappers/test_dead_time_processor.py:28-75 (FIFO protocol: [reset_action] * steps + actions)
  tests/test_physical_systems/test_voltage_supplies.py:86-97 (RC supply system equation), :141-163 (AC1PhaseSupply.get_voltage)
  tests/test_physical_systems/test_mechanical_loads.py:36-42, :281-290 (ExternalSpeedLoad.mechanical_ode, sawtooth profile)
  tests/test_reference_generators/test_reference_generators.py:540-627 (Sawtooth / Sinusoidal / Step / Triangular sub-episodes with the
      tests' fixed random draws: uniform -> 0.25, triangular -> 0.45, testing_utils.py:586-606)
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


def _literals_of(func, names, after=None):
    """the LAST literal assigned to each of `names` in the body of test function `func` (optionally only assignments after the source line
    containing `after`), evaluated with numpy in scope — the tests keep their tables as local literals"""
    import ast
    import inspect
    import textwrap

    src = textwrap.dedent(inspect.getsource(func))
    start = 0
    if after is not None:
        start = next(i for i, l in enumerate(src.splitlines(), 1) if after in l)
    found = {}
    for node in ast.walk(ast.parse(src)):
        if isinstance(node, ast.Assign) and len(node.targets) == 1 and isinstance(node.targets[0], ast.Name) and node.targets[0].id in names \
                and node.lineno >= start:
            found[node.targets[0].id] = eval(compile(ast.Expression(node.value), "<literal>", "eval"), {"np": np})
    missing = set(names) - set(found)
    assert not missing, missing
    return found


def finite_b6_interlock():
    """test_converters.py:634-697, second part of test_discrete_b6_bridge: bridge built from tests/conf.py:336 (tau = 2e-4, interlocking 1e-6),
    ten actions, nineteen convert calls against the expected voltage table; replayed against the reference first"""
    lit = _literals_of(tc.test_discrete_b6_bridge, ("actions", "i_ins", "expected_voltage"), after="# set parameter")
    par = tc.cf.converter_parameter
    conv = cv.FiniteB6BridgeConverter(**par)
    conv.reset()
    calls, k = [], 0
    for t, a, i_in in zip(np.arange(len(lit["actions"])) * par["tau"], lit["actions"], lit["i_ins"]):
        for ts in conv.set_action(a, t):
            u = np.array(conv.convert(np.array(i_in), ts))
            conv.i_sup(np.array(i_in))
            assert all(u == lit["expected_voltage"][k])
            calls.append(dict(action=int(a), t_set=float(t), i_in=[float(x[0]) for x in i_in], t_conv=float(ts), expected=[float(v) for v in lit["expected_voltage"][k]]))
            k += 1
    assert 10 < k <= len(lit["expected_voltage"])  # the table holds one row more than the protocol consumes
    return dict(tau=float(par["tau"]), interlocking_time=float(par["interlocking_time"]), calls=calls)


def cont_b6():
    """test_converters.py:700-791 (test_continuous_b6_bridge): default bridge (u = action / 2) and the parametrised one (interlocking 1e-6 at
    tau = 2e-4) against its expected voltage table; replayed against the reference first"""
    lit = _literals_of(tc.test_continuous_b6_bridge, ("actions", "i_ins", "expected_voltages"))
    out = []
    for par, expected in ((dict(), np.asarray(lit["actions"]) / 2), (dict(tc.cf.converter_parameter), lit["expected_voltages"])):
        conv = cv.ContB6BridgeConverter(**par)
        assert all(conv.reset() == -0.5 * np.ones(3))
        calls = []
        for t, a, i_in, exp in zip(np.arange(len(lit["actions"])) * 1e-4, lit["actions"], lit["i_ins"], expected):
            ts = conv.set_action(np.asarray(a).tolist(), t)
            u = conv.convert(np.array(i_in), ts)
            assert all(abs(float(np.ravel(v)[0]) - e) < 1e-9 for v, e in zip(u, exp))
            calls.append(dict(action=[float(v) for v in a], t_set=float(t), i_in=[float(x[0]) for x in i_in], expected=[float(v) for v in exp],
                              reference_result=[float(np.ravel(v)[0]) for v in u]))
        out.append(dict(tau=float(conv._tau), interlocking_time=float(conv._interlocking_time), calls=calls))
    return out


def i_sup_grid():
    """converter.i_sup (the supply current the RC supply integrates; converters.py:165-184, :289-298, :357-368, :425-435, :484-495, :831-839,
    :903-911).  The reference's class-level tests pin it through properties (test_converters.py:929-935, :967-975, :1046-1055, :1081-1086,
    :1091-1107 `i_sup == action * i_out` without interlocking and `|i_sup| <= |i_out|`, :1146-1154, :1419-1427, :1476-1486 sum over the
    half bridges); here the same grids (plus the actions in between) are evaluated by the reference and recorded call by call."""
    out = []
    taus, ils = (1.0, 2.0), (0.0, 0.1, 1.0)  # :1091-1094
    for kind, cls, actions, i_outs in (("1QC", cv.ContOneQuadrantConverter, ([0.0], [0.5], [1.0]), ([-1.0], [1.0], [0.0])),
                                       ("2QC", cv.ContTwoQuadrantConverter, ([0.0], [0.5], [1.0]), ([0.0], [0.1], [-1.0])),
                                       ("4QC", cv.ContFourQuadrantConverter, ([-1.0], [-0.3], [0.0], [0.5], [1.0]), ([0.0], [1.0], [-2.0])),
                                       ("B6", cv.ContB6BridgeConverter, ([1, -1, 0.65], [0.75, -0.95, -0.3], [-0.25, 0.98, -1], [0, 0, 0]),
                                        ([-1, -1, 0], [1, 1, -2], [0, 0, 1]))):
        for tau in taus:
            for il in ils:
                if il >= tau:
                    continue
                conv = cls(tau=tau, interlocking_time=il)
                conv.reset()
                calls = []
                for n, a in enumerate(actions):
                    conv.set_action(list(a), n * tau)
                    for i_out in i_outs:
                        got = float(conv.i_sup(list(i_out)))  # flat [i_a, i_b, i_c] for the bridge (:911)
                        if kind == "2QC":
                            assert abs(got) <= abs(i_out[0]) and (il > 0 or got == a[0] * i_out[0])
                        calls.append(dict(action=[float(v) for v in a], t_set=float(n * tau), i_out=[float(v) for v in i_out], expected=got))
                out.append(dict(kind=kind, finite=0, tau=tau, interlocking_time=il, calls=calls))
    for kind, cls, i_outs in (("1QC", cv.FiniteOneQuadrantConverter, ([-12.0], [12.0])), ("2QC", cv.FiniteTwoQuadrantConverter, ([-1.0], [0.0], [1.0])),
                              ("4QC", cv.FiniteFourQuadrantConverter, ([-1.0], [0.0], [1.0])),
                              ("B6", cv.FiniteB6BridgeConverter, ([-1, -1, 0], [1, 1, -2], [0, 0, 1]))):
        for il in (0.0, 1e-6):
            conv = cls(tau=1e-5, interlocking_time=il)
            conv.reset()
            calls = []
            for n in range(2 * conv.action_space.n):
                a = (n * 3) % conv.action_space.n
                for ts in conv.set_action(a, n * 1e-5):
                    for i_out in i_outs:
                        conv.convert(list(i_out), ts)
                        calls.append(dict(action=int(a), t_set=float(n * 1e-5), t_conv=float(ts), i_out=[float(v) for v in i_out],
                                          expected=float(conv.i_sup(list(i_out)))))
            out.append(dict(kind=kind, finite=1, tau=1e-5, interlocking_time=il, calls=calls))
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


def poly_load_formula():
    """test_load.py:44-91 (test_polynomial_load): default and parametrised load, omega in {-10, 0, 10}, torque in {-3, 0, 5}; the expected
    value is the test's own closed form sign(omega) (c omega^2 + b |omega| + a), replayed against the reference"""
    import tests.test_physical_systems.test_load as tlo

    tlo.test_polynomial_load()
    out = []
    for load in (tl.PolynomialStaticLoad(), tl.PolynomialStaticLoad(load_parameter=tlo.load_parameter["parameter"])):
        load.set_j_rotor(tlo.load_parameter["j_rot_load"])
        for omega in (-10, 0, 10):
            t_l = float(np.sign(omega) * (load._c * omega ** 2 + load._b * abs(omega) + load._a))
            assert load._static_load(omega) == t_l
            for torque in (-3, 0, 5):
                exp = (torque - t_l) / load.j_total
                assert load.mechanical_ode(0, np.array([omega]), torque) == np.array([exp])
                out.append(dict(a=float(load._a), b=float(load._b), c=float(load._c), j_load=float(load.load_parameter["j_load"]),
                                j_rotor=float(tlo.load_parameter["j_rot_load"]), omega=float(omega), torque=float(torque), expected=float(exp)))
    return out


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


def dq_to_abc():
    """the three (dq action, physical state [omega, epsilon, i], abc action) vectors; DummyPhysicalSystem has tau = 1 (testing_utils.py:132-137),
    the PMSM default p = 3, angle advance 0.5 (dq_to_abc_action_processor.py:69, :87-89).  Replayed against the reference first."""
    import tests.test_physical_system_wrappers.test_dq_to_abc_action_processor as td
    import gym_electric_motor as gem

    names, cases = params_of(td.TestDqToAbcActionProcessor.test_simulate, "dq_action")
    out = []
    for c in cases:
        d = dict(zip(names, c))
        ps_ = td.DummyPhysicalSystem(state_names=["omega", "epsilon", "i"])
        ps_.electrical_motor = gem.physical_systems.PermanentMagnetSynchronousMotor()
        proc = gem.physical_system_wrappers.DqToAbcActionProcessor.make("PMSM", physical_system=ps_)
        proc.reset()
        proc._state = d["state"]
        proc.simulate(d["dq_action"])
        assert all(np.isclose(ps_.action, d["abc_action"]))
        out.append(dict(dq_action=[float(v) for v in d["dq_action"]], omega=float(d["state"][0]), epsilon=float(d["state"][1]), tau=float(ps_.tau),
                        p=float(ps_.electrical_motor.motor_parameter["p"]), angle_advance=float(proc._angle_advance),
                        expected=[float(v) for v in d["abc_action"]], reference_result=[float(v) for v in ps_.action]))
    return out


def dead_time():
    """every (steps, action space, action sequence) combination of test_execution; expected = [reset_action] * steps + actions"""
    import tests.test_physical_system_wrappers.test_dead_time_processor as tdt
    import gym_electric_motor as gem

    f = tdt.TestDeadTimeProcessor.test_execution
    _, procs = params_of(f, "unset_processor")
    names, cases = params_of(f, "action_space")
    out = []
    for proc in procs:
        steps = int(proc.dead_time)
        for c in cases:
            d = dict(zip(names, c))
            ps_ = tdt.DummyPhysicalSystem()
            ps_._action_space = d["action_space"]
            pr = gem.physical_system_wrappers.DeadTimeProcessor(steps=steps)
            pr.set_physical_system(ps_)
            pr.reset()
            expected = [d["reset_action"]] * steps + list(d["actions"])
            seen = []
            for i, a in enumerate(d["actions"]):
                pr.simulate(a)
                assert np.all(np.asarray(ps_.action) == np.asarray(expected[i]))
                seen.append(np.asarray(ps_.action, dtype=float).ravel().tolist())
            kind = type(d["action_space"]).__name__
            out.append(dict(steps=steps, space=kind, actions=[np.asarray(a, dtype=float).ravel().tolist() for a in d["actions"]], applied=seen))
    return out


def supplies():
    """RC: the (u_sup, u_0, i_sup, R, C) -> du/dt literals of test_system_equation, parsed from the test's source; AC1: the parameter sets and
    times of test_get_voltage evaluated by the reference (the reference test itself is executed first so the values are known to satisfy it)."""
    import ast
    import inspect
    import textwrap

    import tests.test_physical_systems.test_voltage_supplies as tv

    tv.TestRCVoltageSupply().test_system_equation()
    tv.TestAC1PhaseSupply().test_get_voltage()
    rc = []
    tree = ast.parse(textwrap.dedent(inspect.getsource(tv.TestRCVoltageSupply.test_system_equation)))
    for node in ast.walk(tree):
        if isinstance(node, ast.Assert) and isinstance(node.test, ast.Compare) and isinstance(node.test.left, ast.Call) \
                and not isinstance(node.test.comparators[0], ast.Call):
            t, u, u0, i_sup, r, c = [ast.literal_eval(a) for a in node.test.left.args]
            expected = eval(compile(ast.Expression(node.test.comparators[0]), "<expected>", "eval"))
            assert tv.vs.RCVoltageSupply().system_equation(t, u, u0, i_sup, r, c) == expected
            rc.append(dict(u_sup=float(u[0]), u_0=float(u0), i_sup=float(i_sup), R=float(r), C=float(c), expected=float(expected)))
    assert len(rc) == 3
    ac = []
    for par, times in ((dict(frequency=1, phase=0), [0, 1, 2, 1 / 4, 5 / 4, 9 / 4, 3 / 4, 7 / 4, 11 / 4]),           # :146-158
                       (dict(frequency=36, phase=0.5), [1 / (2 * np.pi), 2 / (2 * np.pi), 3 / (2 * np.pi)])):       # :160-163
        sup = tv.vs.AC1PhaseSupply(supply_parameter=par)
        for t in times:
            ac.append(dict(u_nominal=float(sup.u_nominal), frequency=float(par["frequency"]), phase=float(par["phase"]), t=float(t),
                           expected=float(sup.get_voltage(t)[0])))
    # the hand-calculated literals of :161-163 must be among them
    assert np.allclose([a["expected"] for a in ac[-3:]], [-303.058731, -78.381295, 323.118651])
    return dict(rc=rc, ac1=ac)


def ext_speed_load():
    import tests.test_physical_systems.test_mechanical_loads as tm

    names, cases = params_of(tm.TestExtSpeedLoad.test_mechanical_ode, "expected_result")
    load = tm.ExternalSpeedLoad(speed_profile=tm.speed_profile_, speed_profile_kwargs=dict(amp=tm.test_amp, bias=tm.test_bias, freq=tm.test_freq))
    out = []
    for omega, expected in cases:
        got = load.mechanical_ode(1, np.array([omega]))[0]
        assert abs(got - expected) < 1e-6
        out.append(dict(omega=float(omega), t=1.0, expected=float(expected), reference_result=float(got)))
    return dict(amp=float(tm.test_amp), bias=float(tm.test_bias), freq=float(tm.test_freq), tau_load=float(load._tau), cases=out)


def periodic_references():
    """test_reset_reference of TestFurtherReferenceGenerator: the scenario literals of the test body (:593-597), the tests' DummyRandom draws,
    `_get_current_value` patched to the identity as in the test; replayed against the reference before recording"""
    import tests.test_reference_generators.test_reference_generators as tr
    from gym_electric_motor.reference_generators.subepisoded_reference_generator import SubepisodedReferenceGenerator

    names, cases = params_of(tr.TestFurtherReferenceGenerator.test_reset_reference, "expected_reference")
    rnd = tr.DummyRandom()

    class FixedDraws:
        def uniform(self, mu=0, sigma=1):
            return rnd.monkey_random_rand()

        def triangular(self, left=-1, mode=0, right=1):
            return rnd.monkey_random_triangular(left, mode, right)

    out = []
    orig = SubepisodedReferenceGenerator._get_current_value
    SubepisodedReferenceGenerator._get_current_value = lambda self, value: value
    try:
        for c in cases:
            d = dict(zip(names, c))
            gen = d["reference_class"](amplitude_range=0.8, frequency_range=d["frequency_range"], offset_range=0.5, limit_margin=0.4, episode_lengths=10,
                                       reference_state="dummy_state_0")
            gen._random_generator = FixedDraws()
            ps_ = tr.DummyPhysicalSystem()
            gen.set_modules(ps_)
            gen._reset_reference()
            assert sum(abs(d["expected_reference"] - gen._reference)) < 1e-6
            out.append(dict(kind=d["reference_class"].__name__, amplitude=0.8, frequency=float(d["frequency_range"]), offset=0.5,
                            margin=[float(v) for v in gen._limit_margin], length=10, tau=float(ps_.tau), uniform_draw=float(FixedDraws().uniform()),
                            triangular_draw=float(FixedDraws().triangular(0, 0.5, 1)), expected=[float(v) for v in d["expected_reference"]],
                            reference_result=[float(v) for v in gen._reference]))
    finally:
        SubepisodedReferenceGenerator._get_current_value = orig
    return out


def wiener_walk():
    """TestWienerProcessReferenceGenerator.test_reset_reference (:384-414): clipped cumulative walk from 0.5 with the DummyRandom increments;
    and TestSubepisodedReferenceGenerator.test_get_current_value (:840-870): lo + (hi - lo) * 0.25, numbers pass through"""
    import tests.test_reference_generators.test_reference_generators as tr
    from gym_electric_motor.reference_generators import WienerProcessReferenceGenerator
    from gym_electric_motor.reference_generators.subepisoded_reference_generator import SubepisodedReferenceGenerator

    rnd = tr.DummyRandom(exp_loc=0, exp_scale=1e-2, exp_size=10)  # :392
    rnd_u = tr.DummyRandom()

    class FixedDraws:
        def normal(self, loc=0, scale=1, size=1):
            return rnd.monkey_random_normal(loc, scale, size)

        def uniform(self, mu=0, sigma=1):
            return rnd_u.monkey_random_rand()

    gen = WienerProcessReferenceGenerator(sigma_range=1e-2, episode_lengths=10, limit_margin=(-1, 1))  # literals of the test body :385-388
    gen._random_generator = FixedDraws()
    gen._reference_value = 0.5
    gen._current_episode_length = 10
    gen._limit_margin = (-1, 1)
    orig = SubepisodedReferenceGenerator._get_current_value
    SubepisodedReferenceGenerator._get_current_value = lambda self, value: value  # as the test does (:396-400)
    try:
        gen._reset_reference()
    finally:
        SubepisodedReferenceGenerator._get_current_value = orig
    expected = [0.6, 0.4, 1, 1, 0.5, 0.2, -1, -0.9, -1, -0.6]  # :389
    assert sum(abs(gen._reference - np.array(expected))) < 1e-6
    walk = dict(start=0.5, margin=[-1.0, 1.0], increments=[float(v) for v in rnd.monkey_random_normal(0, 1e-2, 10)], expected=[float(v) for v in expected],
                reference_result=[float(v) for v in gen._reference])
    names, cases = params_of(tr.TestSubepisodedReferenceGenerator.test_get_current_value, "expected_value")
    sub = SubepisodedReferenceGenerator()
    sub._random_generator = FixedDraws()
    cur = []
    for value_range, expected_value in cases:
        got = sub._get_current_value(value_range)
        assert abs(got - expected_value) < 1e-6
        cur.append(dict(value_range=np.asarray(value_range, dtype=float).ravel().tolist(), uniform_draw=0.25, expected=float(expected_value),
                        reference_result=float(got)))
    return dict(walk=walk, current_value=cur)


if __name__ == "__main__":
    ka = dict(i_sup=i_sup_grid(), finite_b6_interlock=finite_b6_interlock(), cont_b6=cont_b6(), wiener_walk=wiener_walk(), poly_load_formula=poly_load_formula(), periodic_references=periodic_references(), dq_to_abc=dq_to_abc(), dead_time=dead_time(), supplies=supplies(), ext_speed_load=ext_speed_load(),
              finite_qc=finite_qc(), finite_b6=finite_b6(), cont_qc=cont_qc(), poly_load=poly_load(), euler=euler_solver(), constraints=constraints(),
              wse_rewards=wse_rewards())
    with open(os.path.join(HERE, "known_answers.json"), "w") as f:
        json.dump(ka, f)
    print({k: (len(v) if isinstance(v, list) else len(v.get("cases", v))) for k, v in ka.items()},
          "convert calls:", sum(len(c["calls"]) for k in ("finite_qc", "finite_b6", "cont_qc", "cont_b6") for c in ka[k]) + len(ka["finite_b6_interlock"]["calls"]))

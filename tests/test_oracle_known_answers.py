"""Pin the CPU oracle to the KNOWN-ANSWER VECTORS of the reference's own unit tests for the hot path.

tests/golden/known_answers.json is produced by tests/golden/make_known_answers.py, which imports the reference's test modules
(tables in module globals / pytest.mark.parametrize arguments), replays the reference tests' call protocol against the reference
implementation, and records every call with the value the reference's TABLE demands.  Here the same calls go through the oracle's
component-level probes: converters (3509 convert calls and the supply current i_sup of every converter kind: finite 1QC/2QC/4QC with and without interlocking over three taus, the B6
bridge leg by leg, continuous 1QC/2QC/4QC against the tests' `comparable_voltage`), PolynomialStaticLoad.mechanical_ode (known answers
and the closed form of test_load.py), ExternalSpeedLoad.mechanical_ode, the RC supply equation and AC1 supply voltages, the
DqToAbcActionProcessor vectors, the DeadTimeProcessor FIFO protocol, one sub-episode of each periodic reference generator and the
Wiener walk with the tests' fixed random draws, `_get_current_value`, the Limit / Squared constraint truth tables and the
WeightedSumOfErrors cases.
"""
import json
import os

import numpy as np
import pytest

from helpers import GOLDEN_DIR
from gym_electric_motor_b200 import _cabi as K

KA = json.load(open(os.path.join(GOLDEN_DIR, "known_answers.json")))
CONV = {"1QC": K.CONV_1QC, "2QC": K.CONV_2QC, "4QC": K.CONV_4QC}


def _base_cfg(motor=K.MOTOR_PERMEX_DC, conv=(K.CONV_4QC, K.CONV_NONE), finite=0, tau=1e-4, il=0.0):
    cfg = K.new_config()
    cfg.n_envs, cfg.motor_kind, cfg.finite, cfg.tau, cfg.interlocking_time, cfg.u_sup = 1, motor, finite, tau, il, 1.0
    cfg.converter_kind[0], cfg.converter_kind[1] = conv
    cfg.load_kind = K.LOAD_CONST_SPEED
    cfg.solver_kind = K.SOLVER_EULER
    for k, v in ((K.MP_R_A, 1.0), (K.MP_L_A, 1.0), (K.MP_PSI_E, 1.0), (K.MP_P, 1.0), (K.MP_R_S, 1.0), (K.MP_L_D, 1.0), (K.MP_L_Q, 1.0), (K.MP_J_ROTOR, 0.0)):
        cfg.motor_param[k] = v
    cfg.load_param[K.LP_J_LOAD] = 1.0
    return cfg


@pytest.mark.parametrize("case", KA["finite_qc"], ids=lambda c: f"{c['kind']}-tau{c['tau']:g}-il{c['interlocking_time']:g}")
def test_finite_quadrant_converters_follow_the_reference_tables(oracle_lib, case):
    """test_converters.py:313-367 with the tables :14-257"""
    sim = oracle_lib.Oracle(_base_cfg(conv=(CONV[case["kind"]], K.CONV_NONE), finite=1, tau=case["tau"], il=case["interlocking_time"]))
    assert sim.probe_conv_reset()[0] == 0.0
    last = None
    for k, c in enumerate(case["calls"]):
        if (c["action"], c["t_set"]) != last:
            sim.probe_set_action(c["action"], c["t_set"])
            last = (c["action"], c["t_set"])
        assert sim.probe_convert([c["i_in"]], c["t_conv"])[0] == c["expected"], (k, c)


@pytest.mark.parametrize("case", KA["finite_b6"], ids=lambda c: f"leg{c['leg']}")
def test_finite_b6_bridge_follows_the_reference_table(oracle_lib, case):
    """test_converters.py:592-640"""
    sim = oracle_lib.Oracle(_base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE), finite=1, tau=case["tau"]))
    assert list(sim.probe_conv_reset()[:3]) == [-0.5, -0.5, -0.5]
    last = None
    for c in case["calls"]:
        if (c["action"], c["t_set"]) != last:
            sim.probe_set_action(c["action"], c["t_set"])
            last = (c["action"], c["t_set"])
        assert sim.probe_convert(c["i_in"], c["t_conv"])[c["leg"]] == c["expected"]


@pytest.mark.parametrize("case", KA["cont_qc"], ids=lambda c: f"{c['kind']}-tau{c['tau']:g}-il{c['interlocking_time']:g}")
def test_continuous_quadrant_converters_follow_the_reference_formula(oracle_lib, case):
    """test_converters.py:419-503 (seed(123) actions, every current of g_i_ins_cont, expected = the test's comparable_voltage)"""
    sim = oracle_lib.Oracle(_base_cfg(conv=(CONV[case["kind"]], K.CONV_NONE), finite=0, tau=case["tau"], il=case["interlocking_time"]))
    last = None
    for c in case["calls"]:
        if (c["action"], c["t_set"]) != last:
            sim.probe_set_action(c["action"], c["t_set"])
            last = (c["action"], c["t_set"])
        assert abs(sim.probe_convert([c["i_in"]], c["t_conv"])[0] - c["expected"]) < 1e-5  # the reference test's own tolerance :468


def test_polynomial_static_load_known_answers(oracle_lib):
    """test_mechanical_loads.py:191-211: omega = -3 / 0 / 5 -> 23400 / 20000 / 11400 (all three branches of the static torque)"""
    pl = KA["poly_load"]
    cfg = _base_cfg()
    cfg.load_kind = K.LOAD_POLY_STATIC
    lp = pl["load_parameter"]
    cfg.load_param[K.LP_A], cfg.load_param[K.LP_B], cfg.load_param[K.LP_C], cfg.load_param[K.LP_J_LOAD] = lp["a"], lp["b"], lp["c"], lp["j_load"]
    cfg.load_param[K.LP_TAU_DECAY] = 1e-3
    sim = oracle_lib.Oracle(cfg)
    for c in pl["cases"]:
        assert abs(sim.probe_mechanical_ode(c["omega"], pl["torque"]) - c["expected"]) < 1e-6  # abs_tol of the reference test


@pytest.mark.parametrize("case", KA["euler"], ids=lambda c: f"nsteps{c['nsteps']}")
def test_euler_solver_known_answers(oracle_lib, case):
    """test_solvers.py:248-269: the oracle's Euler stepping (the code its motor integration runs) on the reference's test system;
    the table value within the reference test's tolerance, the reference's own float result to round-off"""
    got = oracle_lib.probe_euler(case["nsteps"], case["y0"], case["tau"], case["u"])
    assert np.abs(got - np.array(case["expected"])).sum() < 1e-6
    assert np.abs(got - np.array(case["reference_result"])).max() < 1e-13


@pytest.mark.parametrize("case", KA["constraints"], ids=lambda c: f"{c['kind']}-{c['state']}")
def test_constraint_truth_tables(oracle_lib, case):
    """test_limit_constraint.py:33-66, test_squared_constraint.py:25-99"""
    cfg = _base_cfg()
    cfg.n_constraints = 1
    cfg.constraint_kind[0] = K.CONSTRAINT_SQUARED if case["kind"] == "squared" else K.CONSTRAINT_LIMIT
    cfg.constraint_mask[0] = sum(1 << j for j in case["observed"])
    assert oracle_lib.Oracle(cfg).probe_constraints(case["state"]) == case["expected"]


@pytest.mark.parametrize("case", KA["wse_rewards"], ids=lambda c: f"bias{c['bias']:g}-viol{c['violation_degree']:g}")
def test_weighted_sum_of_errors_cases(oracle_lib, case):
    """test_weighted_sum_of_errors.py:150-218"""
    cfg = _base_cfg()
    for j in range(3):
        cfg.reward_weight[j], cfg.reward_power[j], cfg.state_length[j] = case["reward_weights"][j], case["reward_power"][j], case["state_length"][j]
    cfg.reward_bias, cfg.violation_reward = case["bias"], case["violation_reward"]
    assert oracle_lib.Oracle(cfg).probe_reward(case["state"], case["reference"], case["violation_degree"]) == case["expected"]


@pytest.mark.parametrize("case", KA["dq_to_abc"], ids=lambda c: f"dq{c['dq_action']}")
def test_dq_to_abc_action_processor_known_answers(oracle_lib, case):
    """test_dq_to_abc_action_processor.py:27-53: abc action for a dq action at the advanced angle epsilon + 0.5 tau omega p, through the
    oracle's own action-wrapper code (the part of step_one in front of simulate)."""
    cfg = _base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE), tau=case["tau"])
    cfg.motor_param[K.MP_P] = case["p"]
    cfg.action_dq, cfg.angle_advance = 1, case["angle_advance"]
    sim = oracle_lib.Oracle(cfg)
    sim.reset()
    sim.set_ode_state([[case["omega"], 0.0, 0.0, case["epsilon"]]])
    abc = sim.probe_wrap_action(case["dq_action"])
    assert np.allclose(abc, case["expected"])            # the test's own tolerance (np.isclose)
    assert np.allclose(abc, case["reference_result"], rtol=0, atol=1e-14)   # and the reference's actual output


@pytest.mark.parametrize("case", KA["dead_time"], ids=lambda c: f"steps{c['steps']}-{c['space']}-{len(c['actions'][0])}")
def test_dead_time_processor_fifo_protocol(oracle_lib, case):
    """test_dead_time_processor.py:28-75: the inner system sees [reset_action] * steps + actions.  Box(3) -> a PMSM with abc actions,
    Box(1) -> a DC motor, Discrete -> the finite B6 slot, MultiDiscrete -> the two finite slots of an EESM (its first two columns; a
    third finite slot exists in no system of the reference)."""
    width = len(case["actions"][0])
    if case["space"] == "Box":
        cfg = _base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE)) if width == 3 else _base_cfg()
        cols = width
    elif case["space"] == "Discrete":
        cfg, cols = _base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE), finite=1), 1
    else:
        cfg, cols = _base_cfg(motor=K.MOTOR_EESM, conv=(K.CONV_B6, K.CONV_4QC), finite=1), 2
    cfg.dead_time_steps = case["steps"]
    sim = oracle_lib.Oracle(cfg)
    sim.reset()
    for a, applied in zip(case["actions"], case["applied"]):
        got = sim.probe_wrap_action(a[:cols])
        assert list(got) == applied[:cols]


@pytest.mark.parametrize("case", KA["supplies"]["rc"], ids=lambda c: f"u{c['u_sup']:g}")
def test_rc_supply_system_equation_known_answers(oracle_lib, case):
    """test_voltage_supplies.py:86-97 (hand-calculated values of the reference test)"""
    got = oracle_lib.probe_rc_supply_rhs(case["u_sup"], case["u_0"], case["i_sup"], case["R"], case["C"])
    assert abs(got - case["expected"]) <= 1e-14 * max(1.0, abs(case["expected"]))


def test_ac1_supply_voltage_known_answers(oracle_lib):
    """test_voltage_supplies.py:141-163: zero crossings, peaks and the three hand-calculated values"""
    for c in KA["supplies"]["ac1"]:
        got = oracle_lib.probe_ac1_voltage(c["u_nominal"], c["frequency"], c["phase"], c["t"])
        assert abs(got - c["expected"]) < 1e-9, c


def test_external_speed_load_known_answers(oracle_lib):
    """test_mechanical_loads.py:281-290 with the profile of :36-42 (a triangular wave): d omega/dt = (profile(t + tau) - omega) / tau at
    t = 1.  The profile goes through THIS repo's host tabulation (ExternalSpeedLoad.fill_config), the table entry of t = 1 through the
    oracle's mechanical_ode."""
    from scipy import signal

    from gym_electric_motor_b200.physical_systems import ExternalSpeedLoad

    d = KA["ext_speed_load"]
    load = ExternalSpeedLoad(speed_profile=lambda t, amp, freq, bias: amp * signal.sawtooth(2 * np.pi * freq * t, width=0.5) + bias,
                             speed_profile_kwargs=dict(amp=d["amp"], bias=d["bias"], freq=d["freq"]), tau=d["tau_load"])
    cfg = _base_cfg(tau=1e-4)
    cfg.solver_kind, cfg.solver_nsteps = K.SOLVER_RK4, 1
    load.fill_config(cfg)
    assert cfg.load_kind == K.LOAD_EXT_SPEED
    table = load._table
    dt = 1e-4 / 2  # one table entry per RK4 stage time: tau / (2 nsteps)
    sim = oracle_lib.Oracle(cfg)
    for c in d["cases"]:
        j = int(round(c["t"] / dt))
        got = sim.probe_mechanical_ode_ext(c["omega"], 0.0, table[j])
        assert abs(got - c["expected"]) < 1e-6, c          # the test's own tolerance
        assert abs(got - c["reference_result"]) < 1e-6


@pytest.mark.parametrize("case", KA["periodic_references"], ids=lambda c: c["kind"])
def test_periodic_reference_generators_known_answers(oracle_lib, case):
    """test_reference_generators.py:540-627: one sub-episode of the Sawtooth / Sinusoidal / Step / Triangular generator with the tests'
    fixed draws (uniform -> 0.25, triangular -> 0.45).  The scenario goes through THIS repo's host generator classes (scalar ranges,
    the set_modules clipping of amplitude and offset) into the config, and through the oracle's sub-episode formula with Philox words
    chosen to reproduce the draws."""
    import gym_electric_motor_b200 as gem
    from gym_electric_motor_b200 import reference_generators as rg

    cls = getattr(rg, case["kind"])
    gen = cls(amplitude_range=case["amplitude"], frequency_range=case["frequency"], offset_range=case["offset"], limit_margin=0.4,
              episode_lengths=case["length"], reference_state="omega")
    env = gem.make("Cont-SC-PermExDc-v0", reference_generator=gen, tau=case["tau"])
    cfg = env.build_config()
    assert (cfg.ref_margin_lo[0], cfg.ref_margin_hi[0]) == tuple(case["margin"])
    sim = oracle_lib.Oracle(cfg)

    def word(u):  # the 32-bit word whose u01() is u (to 2^-33)
        return int(u * 2.0 ** 32)

    ratio = case["triangular_draw"]  # StepReferenceGenerator: high/low ratio ~ triangular(0, 0.5, 1); the oracle draws it by inverse CDF
    second = word(2 * ratio * ratio if ratio < 0.5 else 1 - 2 * (1 - ratio) ** 2) if case["kind"].startswith("Step") else word(case["uniform_draw"])
    vals, par = sim.periodic_block(0, cfg.ref_kind[0], [0, 0, 0, 0], [word(case["uniform_draw"]), second, 0, 0], case["length"])
    assert np.sum(np.abs(vals - np.asarray(case["expected"]))) < 1e-6          # the test's own criterion
    assert np.max(np.abs(vals - np.asarray(case["reference_result"]))) < 1e-8  # and the reference's actual sub-episode


def test_wiener_walk_known_answer(oracle_lib):
    """test_reference_generators.py:384-414: clipped cumulative walk from 0.5 with the tests' fixed normal draws"""
    d = KA["wiener_walk"]["walk"]
    cfg = _base_cfg()
    cfg.n_ref, cfg.ref_kind[0], cfg.ref_state[0] = 1, K.REF_WIENER, 0
    cfg.ref_margin_lo[0], cfg.ref_margin_hi[0] = d["margin"]
    sim = oracle_lib.Oracle(cfg)
    got = sim.probe_walk(0, d["start"], d["increments"])
    assert np.sum(np.abs(got - np.asarray(d["expected"]))) < 1e-6
    assert np.array_equal(got, np.asarray(d["reference_result"]))


@pytest.mark.parametrize("case", KA["wiener_walk"]["current_value"], ids=lambda c: str(c["value_range"]))
def test_get_current_value_known_answers(oracle_lib, case):
    """test_reference_generators.py:840-870 (_get_current_value: a number is itself, a range is lo + (hi - lo) * U): through the host
    generator's range handling into the config and the oracle's draw of a sub-episode parameter (the frequency of a periodic generator)"""
    import gym_electric_motor_b200 as gem
    from gym_electric_motor_b200 import reference_generators as rg

    vr = case["value_range"]
    gen = rg.SinusoidalReferenceGenerator(frequency_range=vr[0] if len(vr) == 1 else tuple(vr), reference_state="omega")
    cfg = gem.make("Cont-SC-PermExDc-v0", reference_generator=gen).build_config()
    sim = oracle_lib.Oracle(cfg)
    _, par = sim.periodic_block(0, cfg.ref_kind[0], [0, 0, int(case["uniform_draw"] * 2.0 ** 32), 0], [0, 0, 0, 0], 4)
    assert abs(par[1] - case["expected"]) < 1e-6
    assert abs(par[1] - case["reference_result"]) < 1e-9


def test_polynomial_static_load_closed_form(oracle_lib):
    """test_load.py:44-91: default and parametrised PolynomialStaticLoad, omega in {-10, 0, 10} x torque in {-3, 0, 5}; the reference test
    demands exact equality with (T - sign(omega)(c omega^2 + b |omega| + a)) / j_total"""
    for c in KA["poly_load_formula"]:
        cfg = _base_cfg()
        cfg.load_kind = K.LOAD_POLY_STATIC
        cfg.load_param[K.LP_A], cfg.load_param[K.LP_B], cfg.load_param[K.LP_C], cfg.load_param[K.LP_J_LOAD] = c["a"], c["b"], c["c"], c["j_load"]
        cfg.load_param[K.LP_TAU_DECAY] = 1e-3
        cfg.motor_param[K.MP_J_ROTOR] = c["j_rotor"]
        sim = oracle_lib.Oracle(cfg)
        got = sim.probe_mechanical_ode(c["omega"], c["torque"])
        assert abs(got - c["expected"]) <= 4e-16 * max(1.0, abs(c["expected"])), c


def test_finite_b6_bridge_with_interlocking_follows_the_reference_table(oracle_lib):
    """test_converters.py:634-697: tau = 2e-4, interlocking 1e-6, ten actions, the whole three-leg voltage vector per convert call"""
    case = KA["finite_b6_interlock"]
    sim = oracle_lib.Oracle(_base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE), finite=1, tau=case["tau"], il=case["interlocking_time"]))
    assert list(sim.probe_conv_reset()[:3]) == [-0.5, -0.5, -0.5]
    last = None
    assert len(case["calls"]) >= 18
    for c in case["calls"]:
        if (c["action"], c["t_set"]) != last:
            sim.probe_set_action(c["action"], c["t_set"])
            last = (c["action"], c["t_set"])
        assert list(sim.probe_convert(c["i_in"], c["t_conv"])[:3]) == c["expected"], c


@pytest.mark.parametrize("case", KA["cont_b6"], ids=lambda c: f"il{c['interlocking_time']:g}")
def test_continuous_b6_bridge_known_answers(oracle_lib, case):
    """test_converters.py:700-791: default bridge (u = action / 2) and the parametrised bridge's expected voltages with interlocking"""
    sim = oracle_lib.Oracle(_base_cfg(motor=K.MOTOR_PMSM, conv=(K.CONV_B6, K.CONV_NONE), tau=case["tau"], il=case["interlocking_time"]))
    assert list(sim.probe_conv_reset()[:3]) == [-0.5, -0.5, -0.5]
    for c in case["calls"]:
        sim.probe_set_action(c["action"], c["t_set"])
        u = sim.probe_convert(c["i_in"], c["t_set"] + case["tau"])[:3]
        assert np.max(np.abs(u - np.asarray(c["expected"]))) < 1e-9, c          # the test's own tolerance
        assert np.max(np.abs(u - np.asarray(c["reference_result"]))) < 1e-14


@pytest.mark.parametrize("case", KA["i_sup"], ids=lambda c: f"{'fin' if c['finite'] else 'cont'}-{c['kind']}-tau{c['tau']:g}-il{c['interlocking_time']:g}")
def test_converter_supply_current_follows_the_reference(oracle_lib, case):
    """converter.i_sup over the grids of the reference's class-level converter tests (test_converters.py:929-1486), values recorded from the
    reference call by call: the quantity the RC supply integrates (and the kernel restates)."""
    kind = K.CONV_B6 if case["kind"] == "B6" else CONV[case["kind"]]
    motor = K.MOTOR_PMSM if case["kind"] == "B6" else K.MOTOR_PERMEX_DC
    sim = oracle_lib.Oracle(_base_cfg(motor=motor, conv=(kind, K.CONV_NONE), finite=case["finite"], tau=case["tau"], il=case["interlocking_time"]))
    sim.probe_conv_reset()
    last = None
    for c in case["calls"]:
        if (c["action"], c["t_set"]) != last:
            sim.probe_set_action(c["action"], c["t_set"])
            last = (c["action"], c["t_set"])
        if case["finite"]:
            sim.probe_convert(c["i_out"], c["t_conv"])
        got = sim.probe_i_sup(c["i_out"])
        assert abs(got - c["expected"]) <= 1e-15 * max(1.0, abs(c["expected"])), c

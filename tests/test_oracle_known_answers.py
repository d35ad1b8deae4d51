"""Pin the CPU oracle to the KNOWN-ANSWER VECTORS of the reference's own unit tests for the hot path.

tests/golden/known_answers.json is produced by tests/golden/make_known_answers.py, which imports the reference's test modules
(tables in module globals / pytest.mark.parametrize arguments), replays the reference tests' call protocol against the reference
implementation, and records every call with the value the reference's TABLE demands.  Here the same calls go through the oracle's
component-level probes: converters (3471 convert calls: finite 1QC/2QC/4QC with and without interlocking over three taus, the B6
bridge leg by leg, continuous 1QC/2QC/4QC against the tests' `comparable_voltage`), PolynomialStaticLoad.mechanical_ode, the Limit /
Squared constraint truth tables and the WeightedSumOfErrors cases.
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

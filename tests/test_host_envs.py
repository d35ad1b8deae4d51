"""Host logic (no GPU): the env registry and the limit / nominal / space / reference-margin / reward derivations must
reproduce what the running reference reported for each of its registered ids (tests/golden/env_table.json)."""
import json
import os

import numpy as np
import pytest

import gym_electric_motor_b200 as gem
from gym_electric_motor_b200 import _cabi as K
from helpers import GOLDEN_DIR, config_from_meta, load_golden

TABLE = json.load(open(os.path.join(GOLDEN_DIR, "env_table.json")))
IDS = sorted(TABLE)  # all 54 registered ids of the reference


def test_registry_covers_reference_ids():
    assert sorted(gem.env_ids()) == IDS and len(IDS) == 54
    with pytest.raises(KeyError):
        gem.make("Cont-CC-Nope-v0")


@pytest.mark.parametrize("env_id", IDS)
def test_defaults_match_reference(env_id):
    e = TABLE[env_id]
    env = gem.make(env_id)
    p = env.physical_system
    assert p.state_names == e["state_names"]
    np.testing.assert_allclose(p.limits, e["limits"], rtol=1e-12)
    np.testing.assert_allclose(p.nominal_state, e["nominal_state"], rtol=1e-12)
    np.testing.assert_allclose(p.state_space.low, e["state_low"])
    np.testing.assert_allclose(p.state_space.high, e["state_high"])
    assert p.tau == e["tau"]
    assert p.supply.u_nominal == e["u_sup"]
    assert type(p.electrical_motor).__name__ == e["motor_class"]
    assert type(p.mechanical_load).__name__ == e["load_class"]
    assert type(p.converter).__name__ == e["converter_class"]
    assert p.mechanical_load.j_total == pytest.approx(e["j_total"], rel=1e-12)
    for k_, v in e["motor_parameter"].items():
        if k_ in p.electrical_motor.motor_parameter:
            assert p.electrical_motor.motor_parameter[k_] == v
    sp = e["action_space"]
    if sp["kind"] == "Box":
        np.testing.assert_allclose(env.action_space.low, sp["low"])
        np.testing.assert_allclose(env.action_space.high, sp["high"])
    elif sp["kind"] == "Discrete":
        assert env.action_space.n == sp["n"]
    else:
        assert list(env.action_space.nvec) == sp["nvec"]
    assert list(env.reference_names) == e["reference_names"]
    np.testing.assert_allclose(env.reference_generator.reference_space.low, e["reference_space"]["low"], rtol=1e-12, atol=1e-15)
    np.testing.assert_allclose(env.reference_generator.reference_space.high, e["reference_space"]["high"], rtol=1e-12)
    np.testing.assert_allclose(env.reward_function._reward_weights, e["reward_weights"], rtol=1e-12)
    assert env.reward_function._violation_reward == pytest.approx(e["violation_reward"], rel=1e-12)
    assert list(env.reward_range) == pytest.approx(e["reward_range"])
    # compiled config: constraints, reference slots, initial observation
    cfg = env.build_config()
    names = e["state_names"]
    assert cfg.n_constraints == len(e["constraints"])
    for ci, con in enumerate(e["constraints"]):
        assert cfg.constraint_kind[ci] == (K.CONSTRAINT_SQUARED if con["kind"] == "SquaredConstraint" else K.CONSTRAINT_LIMIT)
        assert cfg.constraint_mask[ci] == sum(1 << names.index(s) for s in con["states"])
    rg = e["reference_generator"]
    subs = rg.get("sub_generators", [rg])
    assert cfg.n_ref == len(subs)
    for r, s in enumerate(subs):
        assert cfg.ref_kind[r] == K.REF_WIENER and cfg.ref_state[r] == names.index(s["reference_state"])
        assert (cfg.ref_margin_lo[r], cfg.ref_margin_hi[r]) == pytest.approx(tuple(s["limit_margin"]), rel=1e-12, abs=1e-15)
        assert (cfg.ref_init_lo[r], cfg.ref_init_hi[r]) == pytest.approx(tuple(s["initial_range"]), rel=1e-12, abs=1e-15)
        assert (cfg.ref_sigma_lo[r], cfg.ref_sigma_hi[r]) == pytest.approx(tuple(s["sigma_range"]))
        assert (cfg.ref_len_lo[r], cfg.ref_len_hi[r]) == tuple(s["episode_len_range"])
    assert cfg.interlocking_time == e["interlocking_time"]
    assert np.allclose([cfg.init_ode[0]], [e["reset_state"][0] * e["limits"][0]])


def test_config_equals_golden_meta_translation():
    """The product's spec compiler and the independent golden-meta translation in tests/helpers.py must agree."""
    for name, env_id, kw in [
        ("pmsm_cc_rk4", "Cont-CC-PMSM-v0", {}),
        ("pmsm_fin_sc_rk4_interlock", "Finite-SC-PMSM-v0", dict(converter=dict(interlocking_time=1e-6))),
        ("eesm_cc_rk4", "Cont-CC-EESM-v0", {}),
        ("scim_sc_rk4", "Cont-SC-SCIM-v0", {}),
        ("permex_cc_rk4", "Cont-CC-PermExDc-v0", {}),
        ("shunt_cc_rk4", "Cont-CC-ShuntDc-v0", {}),
        ("pmsm_sc_polyload_rk4", "Cont-SC-PMSM-v0", dict(load=dict(load_parameter=dict(a=0.5, b=0.02, c=1e-4, j_load=2e-3)))),
        ("dfim_cc_rk4", "Cont-CC-DFIM-v0", {}),
        ("dfim_fin_cc_rk4", "Finite-CC-DFIM-v0", {}),
        ("pmsm_sc_rc_rk4", "Cont-SC-PMSM-v0", dict(supply=gem.physical_systems.RCVoltageSupply(420.0, dict(R=1.0, C=4e-3)))),
        ("extex_cc_rc_rk4", "Cont-CC-ExtExDc-v0", dict(supply=gem.physical_systems.RCVoltageSupply(60.0, dict(R=0.2, C=2e-3)))),
        ("permex_sc_ac_rk4", "Cont-SC-PermExDc-v0", dict(supply=gem.physical_systems.AC1PhaseSupply(42.0, dict(frequency=50.0, phase=0.7)))),
        ("pmsm_cc_custom_rk4", "Cont-CC-PMSM-v0", dict(motor=dict(motor_parameter=dict(p=4, l_d=0.5e-3, l_q=0.9e-3, r_s=25e-3, psi_p=50e-3),
                                                              motor_initializer=dict(states=dict(i_sq=20.0, i_sd=-10.0, epsilon=1.0))))),
    ]:
        g = load_golden(name)
        ref = config_from_meta(g["meta"], reset_ode=g["reset_ode"], solver="rk4", ref_kind=K.REF_WIENER)
        env = gem.make(env_id, ode_solver=gem.physical_systems.RK4Solver(), **kw)
        cfg = env.build_config()
        for f in ("motor_kind", "finite", "load_kind", "solver_kind", "solver_nsteps", "tau", "interlocking_time", "u_sup", "n_ref",
                  "n_constraints", "reward_bias", "violation_reward", "supply_kind"):
            assert getattr(cfg, f) == pytest.approx(getattr(ref, f)), (name, f)
        for f, n in (("converter_kind", 2), ("motor_param", 16), ("load_param", 5), ("limits", 28), ("init_ode", 8), ("reward_weight", 28),
                     ("state_length", 28), ("constraint_mask", 4), ("ref_state", 4), ("supply_param", 3)):
            a, b = list(getattr(cfg, f))[:n], list(getattr(ref, f))[:n]
            assert a == pytest.approx(b, rel=1e-12), (name, f, a, b)


def test_kwargs_semantics():
    env = gem.make("Cont-CC-PMSM-v0", motor=dict(motor_parameter=dict(r_s=0.05)), tau=5e-5,
                   ode_solver=gem.physical_systems.EulerSolver(nsteps=3), supply=dict(u_nominal=400.0), state_filter=["i_sd", "i_sq"])
    assert env.physical_system.electrical_motor.motor_parameter["r_s"] == 0.05
    assert env.physical_system.tau == 5e-5 and env.physical_system.converter.tau == 5e-5
    assert env.physical_system.limits[-1] == 400.0
    assert env.state_filter == [5, 6]
    # env.limits / state_names / nominal_state are those of the FILTERED observation (reference core.py:169-190, tests/test_core.py:255-278)
    assert env.state_names == ["i_sd", "i_sq"] and list(env.limits) == [400.0, 400.0] and list(env.nominal_state) == [240.0, 240.0]
    assert len(env.physical_system.state_names) == 14
    cfg = env.build_config()
    assert cfg.solver_kind == K.SOLVER_EULER and cfg.solver_nsteps == 3
    with pytest.raises(KeyError):
        gem.make("Cont-CC-PMSM-v0", motor=dict(motor_parameter=dict(nope=1)))  # utils.update_parameter_dict
    with pytest.raises(Exception):
        gem.make("Cont-CC-PMSM-v0", converter="Finite-B6C")  # strings are deprecated in the reference (utils.py:12-13)
    with pytest.raises(Exception):
        gem.make("Cont-CC-PMSM-v0", motor=dict(motor_initializer=dict(states=dict(i_sq=1e4, i_sd=0.0, epsilon=0.0))))
    with pytest.raises(TypeError):
        gem.make("Cont-CC-PMSM-v0", motor=object())
    # default (scipy) solver of the reference maps to RK4 with two sub-steps
    assert gem.make("Cont-CC-PMSM-v0").build_config().solver_nsteps == 2


def test_wrapper_descriptors_compile_like_the_reference_stack():
    """physical_system_wrappers=[...] (reference order semantics) -> kernel flags; compared with the independent translation
    of the wrapper goldens in tests/helpers.py."""
    from gym_electric_motor_b200.physical_system_wrappers import DeadTimeProcessor, DqToAbcActionProcessor

    def build(spec):
        return [DeadTimeProcessor(steps=a) if k == "DeadTime" else DqToAbcActionProcessor.make(a) for k, a in spec]

    for name, env_id in [("pmsm_cc_dq_rk4", "Cont-CC-PMSM-v0"), ("pmsm_sc_dq_dead2_rk4", "Cont-SC-PMSM-v0"),
                         ("pmsm_cc_dead1_outer_dq_rk4", "Cont-CC-PMSM-v0"), ("eesm_cc_dq_rk4", "Cont-CC-EESM-v0"),
                         ("pmsm_fin_cc_dead3_rk4", "Finite-CC-PMSM-v0"), ("permex_cc_dead2_rk4", "Cont-CC-PermExDc-v0")]:
        g = load_golden(name)
        ref = config_from_meta(g["meta"], reset_ode=g["reset_ode"], solver="rk4")
        env = gem.make(env_id, ode_solver=gem.physical_systems.RK4Solver(), physical_system_wrappers=build(g["meta"]["case"]["wrappers"]))
        cfg = env.build_config()
        for f in ("action_dq", "dead_time_steps", "dead_time_outer", "angle_advance"):
            assert getattr(cfg, f) == getattr(ref, f), (name, f)
        assert int(np.prod(getattr(env.action_space, "shape", ()) or (1,))) == g["meta"]["action_dim"]
    with pytest.raises(NotImplementedError):
        gem.make("Finite-CC-PMSM-v0", physical_system_wrappers=[DqToAbcActionProcessor.make("PMSM")])
    with pytest.raises(ValueError):  # 'psi_angle' needs a FluxObserver first: the reference's state_names.index() fails (dq_to_abc_action_processor.py:63)
        gem.make("Cont-CC-SCIM-v0", physical_system_wrappers=[DqToAbcActionProcessor.make("SCIM")])
    # control_space='dq' (physical_systems.py:423-435): same transformation, no angle advance
    ps_ = gem.physical_systems
    sys_ = ps_.SynchronousMotorSystem(control_space="dq", supply=ps_.IdealVoltageSupply(300.0), converter=ps_.ContB6BridgeConverter(),
                                      motor=ps_.PermanentMagnetSynchronousMotor(), load=ps_.ConstantSpeedLoad(omega_fixed=100.0),
                                      ode_solver=ps_.RK4Solver())
    c = sys_.fill_config(K.new_config())
    assert (c.action_dq, c.angle_advance) == (1, 0.0) and sys_.action_space.shape == (2,)


def test_state_vector_wrappers_bookkeeping_matches_reference():
    """CosSinProcessor / FluxObserver / StateNoiseProcessor: names, limits, state space and the kernel's op list must equal what the
    reference's wrappers derived when the goldens were recorded (meta) and the independent translation in tests/helpers.py."""
    psw = gem.physical_system_wrappers

    def build(spec):
        out = []
        for k, a in spec:
            out.append(psw.DeadTimeProcessor(steps=a) if k == "DeadTime" else psw.CosSinProcessor(angle=a[0], remove_angle=bool(a[1])) if k == "CosSin"
                       else psw.FluxObserver() if k == "FluxObserver" else psw.DqToAbcActionProcessor.make(a))
        return out

    for name, env_id in [("pmsm_cc_cossin_rk4", "Cont-CC-PMSM-v0"), ("pmsm_sc_cossin_rm_rk4", "Cont-SC-PMSM-v0"), ("scim_cc_flux_rk4", "Cont-CC-SCIM-v0"),
                         ("scim_cc_flux_dq_rk4", "Cont-CC-SCIM-v0"), ("scim_sc_flux_cossin_dead1_rk4", "Cont-SC-SCIM-v0"),
                         ("dfim_cc_flux_dq_rk4", "Cont-CC-DFIM-v0"), ("dfim_sc_dead1_flux_dq_rk4", "Cont-SC-DFIM-v0")]:
        g = load_golden(name)
        meta = g["meta"]
        ref = config_from_meta(meta, reset_ode=g["reset_ode"], solver="rk4", ref_kind=K.REF_WIENER)
        env = gem.make(env_id, ode_solver=gem.physical_systems.RK4Solver(), physical_system_wrappers=build(meta["case"]["wrappers"]))
        ps_ = env.physical_system
        assert ps_.state_names == meta["state_names"], name
        assert np.allclose(ps_.limits, meta["limits"], rtol=1e-12)
        assert np.allclose(ps_.nominal_state, g["meta"]["nominal_state"], rtol=1e-12)
        assert np.allclose(ps_.state_space.low, meta["state_low"]) and np.allclose(ps_.state_space.high, meta["state_high"])
        assert int(np.prod(env.action_space.shape)) == meta["action_dim"]
        cfg = env.build_config()
        assert cfg.n_state_ops == ref.n_state_ops and cfg.action_dq == ref.action_dq and cfg.angle_advance == ref.angle_advance
        for k in range(cfg.n_state_ops):
            assert cfg.sop_kind[k] == ref.sop_kind[k] and list(cfg.sop_idx[k]) == list(ref.sop_idx[k])
            assert list(cfg.sop_param[k]) == pytest.approx(list(ref.sop_param[k]), rel=1e-12)
        for f, n in (("limits", 28), ("reward_weight", 28), ("state_length", 28), ("constraint_mask", 4), ("ref_state", 4)):
            assert list(getattr(cfg, f))[:n] == pytest.approx(list(getattr(ref, f))[:n], rel=1e-12), (name, f)
        n_state = K.C.c_int32()
        K.load_library().gemb200_query_dims(K.C.byref(cfg), K.C.byref(n_state), None, None, None)
        assert n_state.value == len(meta["state_names"])
    # StateNoiseProcessor: descriptor -> op
    env = gem.make("Cont-SC-PermExDc-v0", physical_system_wrappers=[psw.StateNoiseProcessor(states=["omega", "torque"], random_dist="laplace",
                                                                                           random_kwargs=dict(loc=0.0, scale=0.1))])
    cfg = env.build_config()
    assert (cfg.n_state_ops, cfg.sop_kind[0], cfg.sop_idx[0][0], cfg.sop_mask[0]) == (1, K.SOP_NOISE, K.NOISE_LAPLACE, 0b11)
    assert list(cfg.sop_param[0])[:2] == [0.0, 0.1]
    env = gem.make("Cont-CC-PMSM-v0", physical_system_wrappers=[psw.StateNoiseProcessor(states="all")])
    assert env.build_config().sop_mask[0] == (1 << 14) - 1
    with pytest.raises(AssertionError):
        psw.StateNoiseProcessor(states="all", random_dist="not_a_distribution")
    with pytest.raises(NotImplementedError):
        psw.StateNoiseProcessor(states="all", random_dist="gamma")


def test_random_initialiser_bounds_match_reference():
    """random_init='uniform': the [lower, upper] box per ODE state must be the one the reference samples from — compared with
    the empirical range of 3000 reference resets per env (tests/golden/init_bounds.json, make_golden.py:init_bounds)."""
    ref = json.load(open(os.path.join(GOLDEN_DIR, "init_bounds.json")))
    for env_id, e in ref.items():
        env = gem.make(env_id, motor=dict(motor_initializer=dict(random_init="uniform")),
                       load=dict(load_initializer=dict(random_init="uniform", interval=e["load_interval"])))
        cfg = env.build_config()
        assert cfg.init_random == 1
        n = len(e["min"])
        lo, hi = np.array(list(cfg.init_lo)[:n]), np.array(list(cfg.init_hi)[:n])
        mn, mx, mean = np.array(e["min"]), np.array(e["max"]), np.array(e["mean"])
        span = hi - lo
        assert np.all(mn >= lo - 1e-9) and np.all(mx <= hi + 1e-9), (env_id, lo, hi, mn, mx)
        assert np.all(mn - lo < 0.01 * span) and np.all(hi - mx < 0.01 * span), (env_id, lo, hi, mn, mx)
        assert np.all(np.abs(mean - 0.5 * (lo + hi)) < 0.03 * span)
    # induction motors: constant bounds for currents / angle, the flux bounds are re-derived per reset on the device (gemb200.h: init_im)
    cfg = gem.make("Cont-CC-SCIM-v0", motor=dict(motor_initializer=dict(random_init="uniform"))).build_config()
    assert cfg.init_random == 1 and cfg.init_im_valid == 1
    assert list(cfg.init_lo)[:6] == [100.0, -3.9, -3.9, -1e30, -1e30, -np.pi] and list(cfg.init_hi)[:6] == [100.0, 3.9, 3.9, 1e30, 1e30, np.pi]


def test_gaussian_initialiser_config():
    """random_init='gaussian': mue defaults to the middle of [lower, upper], sigma to 1 (electric_motor.py:246-247); given scalars are
    used for every state; the bounds are those of the uniform case."""
    env = gem.make("Cont-SC-PMSM-v0", motor=dict(motor_initializer=dict(random_init="gaussian")),
                   load=dict(load_initializer=dict(random_init="normal", random_params=(20.0, 15.0), interval=[[-50.0, 120.0]])))
    cfg = env.build_config()
    ref = gem.make("Cont-SC-PMSM-v0", motor=dict(motor_initializer=dict(random_init="uniform")),
                   load=dict(load_initializer=dict(random_init="uniform", interval=[[-50.0, 120.0]]))).build_config()
    assert cfg.init_random == 1 and list(cfg.init_lo)[:4] == list(ref.init_lo)[:4] and list(cfg.init_hi)[:4] == list(ref.init_hi)[:4]
    assert list(cfg.init_dist)[:4] == [1, 1, 1, 1] and list(ref.init_dist)[:4] == [0, 0, 0, 0]
    assert (cfg.init_mu[0], cfg.init_sigma[0]) == (20.0, 15.0)
    for j in (1, 2, 3):
        assert cfg.init_mu[j] == pytest.approx(0.5 * (cfg.init_lo[j] + cfg.init_hi[j])) and cfg.init_sigma[j] == 1.0
    with pytest.raises(NotImplementedError):
        gem.make("Cont-SC-PMSM-v0", motor=dict(motor_initializer=dict(random_init="cauchy")))


def test_switched_reference_generator_config():
    """SwitchedReferenceGenerator -> generator-table entries: the output slot plus one parameter entry per sub-generator, cumulative
    probabilities, super-episode range; reference_space = union of the sub-generators' spaces (switched :49-56)."""
    rg = gem.reference_generators
    subs = [rg.WienerProcessReferenceGenerator(reference_state="omega", sigma_range=(1e-3, 1e-2)), rg.SinusoidalReferenceGenerator(reference_state="omega"),
            rg.ConstReferenceGenerator(reference_state="omega", reference_value=0.3)]
    env = gem.make("Cont-SC-PMSM-v0", reference_generator=rg.SwitchedReferenceGenerator(subs, p=[0.5, 0.3, 0.2], super_episode_length=(50, 200)))
    c = env.build_config()
    assert c.n_ref == 1 and (c.ref_sw_count[0], c.ref_sw_first[0], c.ref_sw_len_lo[0], c.ref_sw_len_hi[0]) == (3, 1, 50, 200)
    assert list(c.ref_kind)[1:4] == [K.REF_WIENER, K.REF_SINUS, K.REF_CONST] and c.ref_value[3] == 0.3
    assert list(c.ref_sw_cdf)[1:4] == pytest.approx([0.5, 0.8, 1.0])
    assert env.reference_names == ["omega"] and env.reference_generator.reference_space.shape == (1,)
    with pytest.raises(AssertionError):
        rg.SwitchedReferenceGenerator([rg.WienerProcessReferenceGenerator(reference_state="omega"), rg.WienerProcessReferenceGenerator(reference_state="torque")])
    c8 = gem.make("Cont-SC-PMSM-v0", reference_generator=rg.SwitchedReferenceGenerator([rg.WienerProcessReferenceGenerator(reference_state="omega")] * 8)).build_config()
    assert (c8.ref_sw_count[0], c8.ref_sw_first[0]) == (8, 1) and list(c8.ref_kind)[1:9] == [K.REF_WIENER] * 8  # 1 output slot + 8 entries <= 12
    with pytest.raises(NotImplementedError):  # 1 output + 12 sub-generators > GEMB200_MAX_REF_ENTRIES = 12
        gem.make("Cont-SC-PMSM-v0", reference_generator=rg.SwitchedReferenceGenerator([rg.WienerProcessReferenceGenerator(reference_state="omega")] * 12)).build_config()


def test_external_speed_load_table():
    """ExternalSpeedLoad: the callable is tabulated on the grid of solver stage times (+ tau_load), vectorised or scalar profiles
    alike, and the initial speed is f(0); the table must equal the independent one tests/helpers.py builds from the golden meta."""
    import ctypes as C

    def profile(t, a, f, o):
        return o + a * np.sin(2 * np.pi * f * t)

    def scalar_profile(t, a, f, o):
        return float(o + a * np.sin(2 * np.pi * f * float(t)))  # float() rejects arrays -> per-sample fallback

    g = load_golden("pmsm_cc_extspeed_rk4")
    ref = config_from_meta(g["meta"], reset_ode=g["reset_ode"], solver="rk4")
    for prof in (profile, scalar_profile):
        load = gem.physical_systems.ExternalSpeedLoad(prof, tau=1e-4, speed_profile_kwargs=dict(a=80.0, f=25.0, o=120.0), horizon_steps=1508)
        env = gem.make("Cont-CC-PMSM-v0", load=load, ode_solver=gem.physical_systems.RK4Solver())
        cfg = env.build_config()
        assert cfg.load_kind == K.LOAD_EXT_SPEED == ref.load_kind and cfg.ext_speed_len == ref.ext_speed_len
        assert cfg.load_param[K.LP_TAU_LOAD] == ref.load_param[K.LP_TAU_LOAD] == 1e-4 and cfg.init_ode[0] == 120.0
        a = np.ctypeslib.as_array(C.cast(cfg.ext_speed_table, C.POINTER(C.c_double)), (cfg.ext_speed_len,))
        b = np.ctypeslib.as_array(C.cast(ref.ext_speed_table, C.POINTER(C.c_double)), (ref.ext_speed_len,))
        np.testing.assert_allclose(a, b, rtol=0, atol=1e-12)
    assert gem.make("Cont-CC-PMSM-v0", load=load, ode_solver=gem.physical_systems.EulerSolver(nsteps=3)).build_config().ext_speed_len == 6 * 1508 + 13


def test_vector_facade_spaces():
    venv = gem.vector.make_vec("Cont-CC-PMSM-v0", num_envs=8, flatten_obs=True)
    assert venv.num_envs == 8 and venv.single_observation_space.shape == (16,) and venv.observation_space.shape == (8, 16)
    assert venv.single_action_space.shape == (3,) and venv.action_space.shape == (8, 3)
    v2 = gem.vector.make_vec("Finite-CC-PMSM-v0", num_envs=4)
    assert v2.single_action_space.n == 8 and len(v2.observation_space.spaces) == 2


def test_utils_follow_the_reference_test_vectors():
    """the cases of the reference's tests/test_utils.py (state_dict_to_state_array :8-30, set_state_array :33-56)"""
    from gym_electric_motor_b200 import utils

    names = ["a", "b", "c", "d"]
    for state_dict, state_array, target in ((dict(A=5, b=12, c=10, d=12), [0, 0, 0, 0], [5, 12, 10, 12]), (dict(a=5, b=12, c=10), [0, 1, 2, 5], [5, 12, 10, 5])):
        utils.state_dict_to_state_array(state_dict, state_array, names)
        assert np.all(np.asarray(state_array) == target)
    with pytest.raises(AssertionError):
        utils.state_dict_to_state_array({"invalid_name": 0, "valid_name": 1}, np.zeros(3), ["valid_name", "a", "b"])
    for input_values, target in ((dict(a=5, b=12, c=10, d=12), [5, 12, 10, 12]), (np.array([1, 2, 3, 4]), [1, 2, 3, 4]), ([1, 2, 3, 4], [1, 2, 3, 4]), (5, [5, 5, 5, 5])):
        assert np.all(utils.set_state_array(input_values, names) == target)
    for input_values, error in (("a", Exception), (np.array([1, 2, 3]), AssertionError), ([1, 2, 3], AssertionError)):
        with pytest.raises(error):
            utils.set_state_array(input_values, names)


@pytest.mark.parametrize("env_id", IDS)
def test_env_class_names_and_index_attributes(env_id):
    """one env class per id with the reference's class name (agents test `type(env) in (envs.ContSpeedControl...Env, ...)`) and the
    SCMLSystem index attributes agents read (physical_systems.py:141-162, :462-485, :594-617, :737-763); both recorded from the reference"""
    import gym_electric_motor_b200.envs as envs

    t = TABLE[env_id]
    env = gem.make(env_id)
    assert type(env).__name__ == t["env_class"] and type(env) is getattr(envs, t["env_class"])
    assert isinstance(env, gem.ElectricMotorEnvironment) and env.unwrapped is env
    assert isinstance(env.visualizations[0], gem.visualization.MotorDashboard)  # the reference's default; inert here
    ps = env.physical_system.unwrapped
    for key, val in t["system_indices"].items():
        assert getattr(ps, key) == val, (key, getattr(ps, key), val)


def test_motor_enum_helper_composes_every_id():
    from gym_electric_motor_b200.envs.motors import ActionType, ControlType, Motor, MotorType

    ids = {Motor(m, c, a).env_id() for m in MotorType for c in ControlType for a in ActionType}
    assert ids == set(gem.env_ids())
    assert Motor(MotorType.PermanentMagnetSynchronousMotor, ControlType.TorqueControl, ActionType.Continuous).env_id() == "Cont-TC-PMSM-v0"
    for m in MotorType:  # the plotted state names are states of the env (minus u_sup, and the wrappers' extras)
        env = gem.make(Motor(m, ControlType.SpeedControl, ActionType.Continuous).env_id())
        assert set(Motor(m, ControlType.SpeedControl, ActionType.Continuous).states()) <= set(env.state_names), m


def test_host_side_transformations_match_the_definitions():
    """Clarke / Park helpers agents call (three_phase_motor.py:18-88, physical_systems.py:326-415)"""
    ps = gem.make("Cont-CC-PMSM-v0").physical_system
    m = ps.electrical_motor
    abc = np.array([0.3, -0.7, 0.4])
    ab = m.t_23(abc)
    assert np.allclose(ab, 2 / 3 * np.array([[1, -0.5, -0.5], [0, np.sqrt(3) / 2, -np.sqrt(3) / 2]]) @ abc)
    assert np.allclose(m.t_32(ab), abc - abc.mean())          # zero-sequence free round trip
    dq = np.array(m.q_inv(ab, 0.83))
    assert np.allclose(m.q(dq, 0.83), ab) and np.isclose(np.hypot(*dq), np.hypot(*ab))
    assert np.allclose(ps.abc_to_dq_space(abc, 0.83 / np.pi, normed_epsilon=True), dq)
    assert np.allclose(ps.dq_to_abc_space(dq, 0.83), abc - abc.mean())
    assert np.allclose(ps.alphabeta_to_dq_space(ab, 0.83), dq) and np.allclose(ps.dq_to_alphabeta_space(dq, 0.83), ab)
    assert np.allclose(ps.abc_to_alphabeta_space(abc), ab) and np.allclose(ps.alphabeta_to_abc_space(ab), abc - abc.mean())
    assert np.allclose(m.q_me(dq, 0.2), m.q(dq, 0.2 * m.motor_parameter["p"]))


def test_reference_generator_and_reward_function_can_be_replaced():
    """reference core.py:132-162 / tests/test_core.py:220-262: new component, reset required; here the device handle is rebuilt from the
    new configuration at the next reset"""
    env = gem.make("Cont-SC-PermExDc-v0")
    assert env.build_config().ref_kind[0] == K.REF_WIENER
    env.reference_generator = gem.reference_generators.SinusoidalReferenceGenerator(reference_state="omega", frequency_range=(5, 5))
    cfg = env.build_config()
    assert cfg.ref_kind[0] == K.REF_SINUS and cfg.ref_freq_lo[0] == 5.0 and env.reference_generator.reference_names == ["omega"]
    env.reward_function = gem.reward_functions.WeightedSumOfErrors(reward_weights=dict(omega=2.0), gamma=0.5)
    cfg = env.build_config()
    assert cfg.reward_weight[0] == 2.0 and env.reward_range == env.reward_function.reward_range
    with pytest.raises(TypeError):
        env.reference_generator = object()
    with pytest.raises(TypeError):
        env.reward_function = lambda *a: 0.0


def test_scalar_env_refuses_field_major_layout():
    """the scalar contract (num_envs=None) returns one row per step, which only exists in the row-per-env layout"""
    with pytest.raises(ValueError, match="layout='soa' needs a batched environment"):
        gem.make("Cont-CC-PMSM-v0", layout="soa")
    assert gem.make("Cont-CC-PMSM-v0", layout="soa", num_envs=4).build_config().layout == K.LAYOUT_SOA

#!/usr/bin/env python
"""Generate the golden fixtures under tests/golden/ by running the UNMODIFIED reference.

Runs only in the build container (needs /root/reference); the fixtures it writes are committed and are what
travels to the GPU box.  Usage:  python tests/golden/make_golden.py [--only NAME]

What is recorded per case (one .npz each):
  actions[K,n_act]     the seeded action sequence fed to env.step
  states[K,n_state]    normalised state returned by env.step              (reference core.py:328-371)
  ode_states[K,n_ode]  raw solver state after each step (debug aid)
  refs_used[K,n_state] reference array the reward of step k was computed against (get_reference)
  ref_next[K,n_ref]    reference observation returned by step k
  rewards[K], terminated[K]
  reset_state[n_state], reset_ode[n_ode]   what physical_system.reset() produced
  meta (json)          limits, nominal, names, component parameters, solver, tau
On termination the harness calls env.reset() exactly like the reference's users must (core.py:341) and goes on.

Also written:
  env_table.json       per registered env id: names, limits, nominal state, spaces, defaults
  ref_data_regen.npz   the reference's own integration golden (tests/integration_tests/ref_data.npz) regenerated
                       here and asserted equal to the reference's file at generation time.
"""
import argparse
import json
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = "/root/reference"
sys.dont_write_bytecode = True
try:
    import gymnasium  # noqa: F401
except ImportError:
    sys.path.insert(0, os.path.join(HERE, "..", "_shims"))
sys.path.insert(0, os.path.join(REF_ROOT, "src"))
warnings.filterwarnings("ignore")

import gym_electric_motor as gem  # noqa: E402
from gym_electric_motor import physical_systems as ps  # noqa: E402
from gym_electric_motor.core import ElectricMotorVisualization  # noqa: E402
from gym_electric_motor.physical_systems.solvers import EulerSolver, OdeSolver, ScipyOdeSolver  # noqa: E402


class NoViz(ElectricMotorVisualization):
    """Explicit no-op visualization so the reference never touches matplotlib."""


class RK4Solver(OdeSolver):
    """Classic RK4 with `nsteps` equal sub-steps as a reference-side OdeSolver PLUGIN (the reference ships none).

    Uses only the reference's OdeSolver interface (solvers.py:4-76) so that the reference's own
    converter/motor/load code produces the right-hand side; this pins the RK4 kernels algorithm-for-algorithm.
    """

    def __init__(self, nsteps=1):
        self._nsteps = nsteps

    def integrate(self, t):
        h = (t - self._t) / self._nsteps
        y = np.array(self._y, dtype=float)
        tc = self._t
        f = self._system_equation
        for _ in range(self._nsteps):
            k1 = np.array(f(tc, y, *self._f_params), dtype=float)
            k2 = np.array(f(tc + 0.5 * h, y + 0.5 * h * k1, *self._f_params), dtype=float)
            k3 = np.array(f(tc + 0.5 * h, y + 0.5 * h * k2, *self._f_params), dtype=float)
            k4 = np.array(f(tc + h, y + h * k3, *self._f_params), dtype=float)
            y = y + h / 6.0 * (k1 + 2 * k2 + 2 * k3 + k4)
            tc += h
        self._y = y
        self._t = t
        return self._y


def sin_profile(t, a, f, o):
    return o + a * np.sin(2 * np.pi * f * t)


def make_solver(name):
    if name == "euler":
        return EulerSolver()
    if name.startswith("euler"):
        return EulerSolver(nsteps=int(name[5:]))
    if name == "rk4":
        return RK4Solver()
    if name.startswith("rk4x"):
        return RK4Solver(nsteps=int(name[4:]))
    if name == "dopri5":
        return ScipyOdeSolver()
    raise ValueError(name)


def action_sequence(env, K, seed, style):
    """Seeded synthetic actions: first half iid uniform, second half piecewise-constant holds (so currents build up)."""
    rng = np.random.default_rng(seed)
    space = env.action_space
    if hasattr(space, "n"):
        a = rng.integers(0, space.n, size=K)
        hold = rng.integers(1, 30, size=K)
        out = a.copy()
        k = K // 2
        while k < K:
            out[k : k + hold[k]] = a[k]
            k += hold[k]
        return out.astype(np.int64)
    if hasattr(space, "nvec"):
        a = (rng.random((K, len(space.nvec))) * space.nvec).astype(np.int64)
        return a
    n = space.shape[0]
    low, high = space.low, space.high
    a = rng.uniform(low, high, size=(K, n))
    if style == "iid":
        return a
    amp = rng.uniform(0.02, 1.0, size=K)
    hold = rng.integers(2, 60, size=K)
    out = a.copy()
    k = K // 2
    while k < K:
        mid = 0.5 * (low + high)
        out[k : k + hold[k]] = mid + amp[k] * (a[k] - mid)
        k += hold[k]
    return out


def ode_state(env):
    return np.array(env.physical_system._ode_solver.y, dtype=float)


def describe(env, case):
    p = env.physical_system
    load = p.mechanical_load
    conv = p.converter
    meta = dict(
        case=case,
        action_dim=int(np.prod(getattr(env.action_space, "shape", ()) or (1,))),
        state_names=list(p.state_names),
        base_state_names=list(p.unwrapped.state_names),  # the inner system's own vector (before state-vector wrappers)
        base_limits=np.asarray(p.unwrapped.limits, dtype=float).tolist(),
        limits=np.asarray(p.limits, dtype=float).tolist(),
        nominal_state=p.nominal_state.tolist(),
        state_low=p.state_space.low.tolist(),
        state_high=p.state_space.high.tolist(),
        tau=p.tau,
        motor_class=type(p.electrical_motor).__name__,
        motor_parameter={k: float(v) for k, v in p.electrical_motor.motor_parameter.items()},
        load_class=type(load).__name__,
        j_total=float(load.j_total),
        load_parameter={k: float(v) for k, v in getattr(load, "load_parameter", {}).items()},
        omega_fixed=float(getattr(load, "omega_fixed", 0.0) or 0.0),
        ext_speed=dict(tau=float(getattr(load, "_tau", 0.0)), **{k: float(v) for k, v in getattr(load, "speed_profile_kwargs", {}).items()}),
        supply_class=type(p.supply).__name__,
        u_sup=float(p.supply.u_nominal),
        supply_parameter=dict(R=float(getattr(p.supply, "_r", 0.0)), C=float(getattr(p.supply, "_c", 0.0)), f=float(getattr(p.supply, "_f", 0.0)),
                              phase=float(getattr(p.supply, "_phi", 0.0))),
        converter_class=type(conv).__name__,
        # a multi converter keeps its own (unused) copy; the sub-converters' value is the one in force
        interlocking_time=float(max([conv._interlocking_time] + [sc._interlocking_time for sc in getattr(conv, "_sub_converters", [])])
                                if case.get("multi") is not None else conv._interlocking_time),
        # per sub-converter (slot) when they differ: [slot 0, slot 1]
        interlocking_times=[float(sc._interlocking_time) for sc in getattr(conv, "_sub_converters", [])],
        reference_names=list(env.reference_generator.reference_names),
        referenced_states=np.asarray(env.reference_generator.referenced_states).astype(int).tolist(),
        reward_weights=np.asarray(env.reward_function._reward_weights, dtype=float).tolist(),
        reward_power=np.asarray(env.reward_function._n, dtype=float).tolist(),
        state_length=np.asarray(env.reward_function._state_length, dtype=float).tolist(),
        reward_bias=float(env.reward_function._bias),
        violation_reward=float(env.reward_function._violation_reward),
        constraints=[
            dict(kind=type(c).__name__, states=list(getattr(c, "_states", None) or np.asarray(p.state_names)[c._observed_states]))
            for c in env.constraint_monitor.constraints
        ],
    )
    return meta


def record(case):
    kwargs = dict(case.get("kwargs", {}))
    conv_kw = case.get("converter")
    if conv_kw is not None:
        kwargs["converter"] = dict(conv_kw)
    if case.get("load") is not None:
        kwargs["load"] = dict(case["load"])
    if case.get("motor") is not None:
        kwargs["motor"] = dict(case["motor"])
    if case.get("tau") is not None:
        kwargs["tau"] = case["tau"]
    if case.get("converter_cls") is not None:
        kwargs["converter"] = getattr(ps, case["converter_cls"])(**case.get("converter_args", {}))
    if case.get("supply_rc") is not None:  # [u_nominal, R, C]
        u0, r, c = case["supply_rc"]
        kwargs["supply"] = ps.RCVoltageSupply(u_nominal=u0, supply_parameter=dict(R=r, C=c))
    if case.get("ext_speed") is not None:  # ExternalSpeedLoad with omega(t) = o + a sin(2 pi f t); [a, f, o]
        a_, f_, o_ = case["ext_speed"]
        kwargs["load"] = ps.ExternalSpeedLoad(speed_profile=sin_profile, tau=case.get("tau") or (1e-5 if case["env_id"].startswith("Finite") else 1e-4),
                                              speed_profile_kwargs=dict(a=a_, f=f_, o=o_))
    if case.get("supply_ac") is not None:  # [u_nominal, f, phase]
        u0, f, ph = case["supply_ac"]
        kwargs["supply"] = ps.AC1PhaseSupply(u_nominal=u0, supply_parameter=dict(frequency=f, phase=ph))
    if case.get("multi") is not None:  # multi converter built from INSTANCES, as the reference's env defaults do
        subs = [getattr(ps, c)(**a) for c, a in case["multi"]]
        kwargs["converter"] = (ps.FiniteMultiConverter if case["env_id"].startswith("Finite") else ps.ContMultiConverter)(subconverters=subs)
    if case.get("wrappers"):
        from gym_electric_motor.physical_system_wrappers import (CosSinProcessor, DeadTimeProcessor, DqToAbcActionProcessor,
                                                                 FluxObserver)

        ws = []
        for kind, arg in case["wrappers"]:
            if kind == "DeadTime":
                ws.append(DeadTimeProcessor(steps=arg))
            elif kind == "CosSin":  # arg = [angle name, remove_angle]
                ws.append(CosSinProcessor(angle=arg[0], remove_angle=bool(arg[1])))
            elif kind == "FluxObserver":
                ws.append(FluxObserver())
            else:
                ws.append(DqToAbcActionProcessor.make(arg))
        kwargs["physical_system_wrappers"] = ws
    env = gem.make(case["env_id"], visualization=NoViz(), ode_solver=make_solver(case["solver"]), **kwargs)
    K = case["steps"]
    actions = action_sequence(env, K, case["seed"], case.get("style", "mixed"))
    (s0, _), _ = env.reset(seed=case["seed"])
    n_state = len(env.physical_system.state_names)
    reset_state = np.array(s0, dtype=float)
    reset_ode = ode_state(env)
    states = np.zeros((K, n_state))
    ode_states = np.zeros((K, len(reset_ode)))
    refs_used = np.zeros((K, n_state))
    ref_next = np.zeros((K, len(env.reference_generator.reference_names)))
    rewards = np.zeros(K)
    terminated = np.zeros(K, dtype=np.uint8)
    rg = env.reference_generator
    for k in range(K):
        refs_used[k] = rg.get_reference(None)
        a = actions[k]
        if actions.ndim == 1:
            a = int(a)
        (s, rn), r, term, trunc, _ = env.step(a)
        states[k] = s
        ode_states[k] = ode_state(env)
        ref_next[k] = rn
        rewards[k] = r
        terminated[k] = term
        if term:
            env.reset()
            assert np.allclose(ode_state(env), reset_ode), "golden harness assumes deterministic initial states"
    meta = describe(env, case)
    out = os.path.join(HERE, case["name"] + ".npz")
    np.savez_compressed(
        out,
        actions=actions,
        states=states,
        ode_states=ode_states,
        refs_used=refs_used,
        ref_next=ref_next,
        rewards=rewards,
        terminated=terminated,
        reset_state=reset_state,
        reset_ode=reset_ode,
        meta=json.dumps(meta),
    )
    print(f"{case['name']:42s} K={K} terminations={int(terminated.sum())} max|state|={np.abs(states).max():.3f}")


def C(name, env_id, solver, steps=1500, seed=7, **kw):
    return dict(name=name, env_id=env_id, solver=solver, steps=steps, seed=seed, **kw)


CASES = [
    # BASELINE.json configs[0]: the CPU reference-parity case, 10k random-action steps, Euler.
    C("permex_cc_euler_10k", "Cont-CC-PermExDc-v0", "euler", steps=10000, seed=0, style="iid"),
    C("permex_cc_rk4", "Cont-CC-PermExDc-v0", "rk4", steps=1500),
    C("permex_cc_dopri5", "Cont-CC-PermExDc-v0", "dopri5", steps=1500),
    C("permex_sc_dopri5", "Cont-SC-PermExDc-v0", "dopri5", steps=1500),
    C("permex_sc_euler3", "Cont-SC-PermExDc-v0", "euler3", steps=1500),
    C("permex_fin_cc_rk4", "Finite-CC-PermExDc-v0", "rk4", steps=1500),
    # PMSM (configs[1], [2])
    C("pmsm_cc_euler", "Cont-CC-PMSM-v0", "euler", steps=2000),
    C("pmsm_cc_euler3", "Cont-CC-PMSM-v0", "euler3", steps=1500),
    C("pmsm_cc_rk4", "Cont-CC-PMSM-v0", "rk4", steps=2000),
    C("pmsm_cc_rk4x2", "Cont-CC-PMSM-v0", "rk4x2", steps=1500),
    C("pmsm_cc_dopri5", "Cont-CC-PMSM-v0", "dopri5", steps=2000),
    C("pmsm_cc_rk4_interlock", "Cont-CC-PMSM-v0", "rk4", steps=1500, converter=dict(interlocking_time=2e-6)),
    C("pmsm_tc_rk4", "Cont-TC-PMSM-v0", "rk4", steps=1000),
    C("pmsm_sc_rk4", "Cont-SC-PMSM-v0", "rk4", steps=2000),
    C("pmsm_sc_dopri5", "Cont-SC-PMSM-v0", "dopri5", steps=1500),
    C("pmsm_fin_sc_rk4", "Finite-SC-PMSM-v0", "rk4", steps=3000),
    C("pmsm_fin_sc_dopri5", "Finite-SC-PMSM-v0", "dopri5", steps=2000),
    C("pmsm_fin_sc_euler", "Finite-SC-PMSM-v0", "euler", steps=2000),
    C("pmsm_fin_sc_rk4_interlock", "Finite-SC-PMSM-v0", "rk4", steps=3000, converter=dict(interlocking_time=1e-6)),
    C("pmsm_fin_cc_rk4", "Finite-CC-PMSM-v0", "rk4", steps=2000),
    # SynRM / EESM / SCIM (configs[3], [4])
    C("synrm_cc_rk4", "Cont-CC-SynRM-v0", "rk4", steps=2000),
    C("synrm_cc_dopri5", "Cont-CC-SynRM-v0", "dopri5", steps=1500),
    C("synrm_fin_sc_rk4", "Finite-SC-SynRM-v0", "rk4", steps=2000),
    C("eesm_cc_rk4", "Cont-CC-EESM-v0", "rk4", steps=2000),
    C("eesm_cc_euler", "Cont-CC-EESM-v0", "euler", steps=1500),
    C("eesm_cc_dopri5", "Cont-CC-EESM-v0", "dopri5", steps=1500),
    C("eesm_sc_rk4", "Cont-SC-EESM-v0", "rk4", steps=1500),
    C("eesm_fin_cc_rk4", "Finite-CC-EESM-v0", "rk4", steps=2000),
    C("scim_cc_rk4", "Cont-CC-SCIM-v0", "rk4", steps=2000),
    C("scim_cc_euler", "Cont-CC-SCIM-v0", "euler", steps=1500),
    C("scim_cc_dopri5", "Cont-CC-SCIM-v0", "dopri5", steps=1500),
    C("scim_sc_rk4", "Cont-SC-SCIM-v0", "rk4", steps=1500),
    C("scim_fin_sc_rk4", "Finite-SC-SCIM-v0", "rk4", steps=2000),
    # doubly fed induction motor: two B6 bridges (stator, rotor)
    C("dfim_cc_rk4", "Cont-CC-DFIM-v0", "rk4", steps=2000),
    C("dfim_cc_euler", "Cont-CC-DFIM-v0", "euler", steps=1500),
    C("dfim_cc_dopri5", "Cont-CC-DFIM-v0", "dopri5", steps=1500),
    C("dfim_sc_rk4", "Cont-SC-DFIM-v0", "rk4", steps=1500),
    C("dfim_cc_interlock_rk4", "Cont-CC-DFIM-v0", "rk4", steps=1500,
      multi=[("ContB6BridgeConverter", dict(interlocking_time=2e-6)), ("ContB6BridgeConverter", dict(interlocking_time=2e-6))]),
    C("dfim_fin_cc_rk4", "Finite-CC-DFIM-v0", "rk4", steps=2000),
    C("dfim_fin_sc_interlock_rk4", "Finite-SC-DFIM-v0", "rk4", steps=2000,
      multi=[("FiniteB6BridgeConverter", dict(interlocking_time=1e-6)), ("FiniteB6BridgeConverter", dict(interlocking_time=1e-6))]),
    # RC voltage supply (voltage_supplies.py:75-123) behind every converter family
    C("permex_sc_rc_rk4", "Cont-SC-PermExDc-v0", "rk4", steps=1500, supply_rc=[60.0, 0.5, 4e-3]),
    C("permex_fin_sc_rc_interlock_rk4", "Finite-SC-PermExDc-v0", "rk4", steps=2000, supply_rc=[60.0, 0.5, 4e-3], converter=dict(interlocking_time=1e-6)),
    C("extex_cc_rc_rk4", "Cont-CC-ExtExDc-v0", "rk4", steps=1500, supply_rc=[60.0, 0.2, 2e-3]),
    C("pmsm_sc_rc_rk4", "Cont-SC-PMSM-v0", "rk4", steps=1500, supply_rc=[420.0, 1.0, 4e-3]),
    C("pmsm_fin_cc_rc_rk4", "Finite-CC-PMSM-v0", "rk4", steps=2000, supply_rc=[420.0, 1.0, 1e-3]),
    C("pmsm_cc_rc_interlock_euler3", "Cont-CC-PMSM-v0", "euler3", steps=1500, supply_rc=[300.0, 0.3, 4e-3], converter=dict(interlocking_time=2e-6)),
    C("eesm_fin_cc_rc_rk4", "Finite-CC-EESM-v0", "rk4", steps=2000, supply_rc=[300.0, 1.0, 2e-3]),
    C("dfim_cc_rc_rk4", "Cont-CC-DFIM-v0", "rk4", steps=1500, supply_rc=[420.0, 1.0, 4e-3]),
    # ExternalSpeedLoad (external_speed_load.py): sinusoidal speed profiles on the fixed-step solvers
    C("pmsm_cc_extspeed_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1500, ext_speed=[80.0, 25.0, 120.0]),
    C("permex_cc_extspeed_euler3", "Cont-CC-PermExDc-v0", "euler3", steps=1500, ext_speed=[150.0, 40.0, 50.0]),
    C("scim_cc_extspeed_rk4x2", "Cont-CC-SCIM-v0", "rk4x2", steps=1500, ext_speed=[100.0, 10.0, 150.0]),
    C("pmsm_fin_cc_extspeed_euler", "Finite-CC-PMSM-v0", "euler", steps=2000, ext_speed=[200.0, 100.0, 0.0]),
    # single-phase AC supply with a fixed phase (voltage_supplies.py:126-166)
    C("permex_sc_ac_rk4", "Cont-SC-PermExDc-v0", "rk4", steps=1500, supply_ac=[42.0, 50.0, 0.7]),
    C("series_fin_cc_ac_interlock_rk4", "Finite-CC-SeriesDc-v0", "rk4", steps=2000, supply_ac=[230.0, 400.0, 2.5], converter=dict(interlocking_time=1e-6)),
    C("pmsm_cc_ac_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1500, supply_ac=[230.0, 50.0, 4.0]),
    # cross-feature combinations: supply x wrappers x load x converter family
    C("eesm_cc_rc_dq_dead1_rk4", "Cont-CC-EESM-v0", "rk4", steps=1500, supply_rc=[300.0, 0.5, 2e-3], wrappers=[("DeadTime", 1), ("DqToAbc", "EESM")]),
    C("dfim_fin_cc_rc_rk4", "Finite-CC-DFIM-v0", "rk4", steps=2000, supply_rc=[420.0, 1.0, 1e-3]),
    C("extex_fin_cc_ac_rk4", "Finite-CC-ExtExDc-v0", "rk4", steps=2000, supply_ac=[42.0, 400.0, 1.0]),
    C("scim_cc_extspeed_flux_dq_rk4", "Cont-CC-SCIM-v0", "rk4", steps=1500, ext_speed=[60.0, 15.0, 120.0],
      wrappers=[("FluxObserver", None), ("DqToAbc", "SCIM")]),
    C("pmsm_sc_rc_cossin_dq_euler3", "Cont-SC-PMSM-v0", "euler3", steps=1500, supply_rc=[420.0, 0.5, 4e-3],
      wrappers=[("CosSin", ["epsilon", 0]), ("DqToAbc", "PMSM")]),
    # remaining DC family
    C("series_cc_rk4", "Cont-CC-SeriesDc-v0", "rk4", steps=1500),
    C("series_sc_dopri5", "Cont-SC-SeriesDc-v0", "dopri5", steps=1500),
    C("extex_cc_rk4", "Cont-CC-ExtExDc-v0", "rk4", steps=1500),
    C("extex_sc_dopri5", "Cont-SC-ExtExDc-v0", "dopri5", steps=1500),
    C("extex_fin_cc_rk4", "Finite-CC-ExtExDc-v0", "rk4", steps=1500),
    C("shunt_cc_rk4", "Cont-CC-ShuntDc-v0", "rk4", steps=1500),
    C("shunt_fin_sc_rk4", "Finite-SC-ShuntDc-v0", "rk4", steps=1500),
    # elementary converters (converters.py:218-310, :371-435) incl. interlocking
    C("permex_cont1qc_rk4", "Cont-CC-PermExDc-v0", "rk4", steps=1000, converter_cls="ContOneQuadrantConverter"),
    C("permex_cont2qc_rk4", "Cont-CC-PermExDc-v0", "rk4", steps=1000, converter_cls="ContTwoQuadrantConverter",
      converter_args=dict(interlocking_time=3e-6)),
    C("permex_cont4qc_interlock_rk4", "Cont-SC-PermExDc-v0", "rk4", steps=1000, converter=dict(interlocking_time=5e-6)),
    C("permex_fin1qc_rk4", "Finite-CC-PermExDc-v0", "rk4", steps=1000, converter_cls="FiniteOneQuadrantConverter"),
    C("permex_fin2qc_interlock_rk4", "Finite-CC-PermExDc-v0", "rk4", steps=1000, converter_cls="FiniteTwoQuadrantConverter",
      converter_args=dict(interlocking_time=1e-6)),
    C("permex_fin4qc_interlock_rk4", "Finite-SC-PermExDc-v0", "rk4", steps=1000, converter=dict(interlocking_time=1e-6)),
    C("scim_fin_cc_interlock_rk4", "Finite-CC-SCIM-v0", "rk4", steps=1500, converter=dict(interlocking_time=1e-6)),
    # multi converters whose sub-converters have DIFFERENT interlocking times (converters.py:498-740): per-slot dead time of the continuous
    # converters; finite: up to three switching segments per step, the legs of the sub-converter with the shorter time reach their commanded
    # state at the other one's switching time (converters.py:273)
    C("extex_cc_interlock2_rk4", "Cont-CC-ExtExDc-v0", "rk4", steps=1500,
      multi=[("ContFourQuadrantConverter", dict(interlocking_time=1e-6)), ("ContTwoQuadrantConverter", dict(interlocking_time=3e-6))]),
    C("eesm_cc_interlock2_rk4", "Cont-CC-EESM-v0", "rk4", steps=1500,
      multi=[("ContB6BridgeConverter", dict(interlocking_time=2e-6)), ("ContFourQuadrantConverter", dict(interlocking_time=5e-6))]),
    C("extex_fin_cc_interlock2_rk4", "Finite-CC-ExtExDc-v0", "rk4", steps=2500,
      multi=[("FiniteFourQuadrantConverter", dict(interlocking_time=1e-6)), ("FiniteFourQuadrantConverter", dict(interlocking_time=2.5e-6))]),
    C("extex_fin_cc_interlock2b_rk4", "Finite-CC-ExtExDc-v0", "rk4", steps=2500,
      multi=[("FiniteFourQuadrantConverter", dict(interlocking_time=3e-6)), ("FiniteTwoQuadrantConverter", dict())]),
    C("dfim_fin_sc_interlock2_rk4", "Finite-SC-DFIM-v0", "rk4", steps=2000,
      multi=[("FiniteB6BridgeConverter", dict(interlocking_time=2e-6)), ("FiniteB6BridgeConverter", dict(interlocking_time=0.5e-6))]),
    # non-default parameters: load polynomial with all terms, custom motor, non-zero constant initial state
    C("pmsm_sc_polyload_rk4", "Cont-SC-PMSM-v0", "rk4", steps=1500,
      load=dict(load_parameter=dict(a=0.5, b=0.02, c=1e-4, j_load=2e-3))),
    # physical-system wrappers: dq actions with angle advance, dead time, both orders
    C("pmsm_cc_dq_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1500, wrappers=[("DqToAbc", "PMSM")]),
    C("pmsm_sc_dq_dead2_rk4", "Cont-SC-PMSM-v0", "rk4", steps=1500, wrappers=[("DeadTime", 2), ("DqToAbc", "PMSM")]),
    C("pmsm_cc_dead1_outer_dq_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1500, wrappers=[("DqToAbc", "PMSM"), ("DeadTime", 1)]),
    C("eesm_cc_dq_rk4", "Cont-CC-EESM-v0", "rk4", steps=1500, wrappers=[("DqToAbc", "EESM")]),
    C("pmsm_fin_cc_dead3_rk4", "Finite-CC-PMSM-v0", "rk4", steps=1500, wrappers=[("DeadTime", 3)]),
    C("permex_cc_dead2_rk4", "Cont-CC-PermExDc-v0", "rk4", steps=1000, wrappers=[("DeadTime", 2)]),
    # state-vector wrappers: CosSinProcessor (with and without removing the angle), FluxObserver, FluxObserver + dq actions (SCIM)
    C("pmsm_cc_cossin_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1000, wrappers=[("CosSin", ["epsilon", 0])]),
    C("pmsm_sc_cossin_rm_rk4", "Cont-SC-PMSM-v0", "rk4", steps=1000, wrappers=[("CosSin", ["epsilon", 1])]),
    C("scim_cc_flux_rk4", "Cont-CC-SCIM-v0", "rk4", steps=1500, wrappers=[("FluxObserver", None)]),
    C("scim_cc_flux_dq_rk4", "Cont-CC-SCIM-v0", "rk4", steps=1500, wrappers=[("FluxObserver", None), ("DqToAbc", "SCIM")]),
    C("scim_sc_flux_cossin_dead1_rk4", "Cont-SC-SCIM-v0", "rk4", steps=1500,
      wrappers=[("DeadTime", 1), ("FluxObserver", None), ("CosSin", ["psi_angle", 0]), ("DqToAbc", "SCIM")]),
    C("dfim_cc_flux_dq_rk4", "Cont-CC-DFIM-v0", "rk4", steps=1500, wrappers=[("FluxObserver", None), ("DqToAbc", "DFIM")]),
    C("dfim_sc_dead1_flux_dq_rk4", "Cont-SC-DFIM-v0", "rk4", steps=1500, wrappers=[("DeadTime", 1), ("FluxObserver", None), ("DqToAbc", "DFIM")]),
    C("pmsm_cc_custom_rk4", "Cont-CC-PMSM-v0", "rk4", steps=1500,
      motor=dict(motor_parameter=dict(p=4, l_d=0.5e-3, l_q=0.9e-3, r_s=25e-3, psi_p=50e-3),
                 motor_initializer=dict(states=dict(i_sq=20.0, i_sd=-10.0, epsilon=1.0)))),
]


def jsonable(x):
    if isinstance(x, dict):
        return {str(k): jsonable(v) for k, v in x.items()}
    if isinstance(x, (list, tuple)):
        return [jsonable(v) for v in x]
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, (int, float, str, bool)) or x is None:
        return x
    return repr(x)


def space_desc(sp):
    if hasattr(sp, "nvec"):
        return dict(kind="MultiDiscrete", nvec=np.asarray(sp.nvec).tolist())
    if hasattr(sp, "n"):
        return dict(kind="Discrete", n=int(sp.n))
    return dict(kind="Box", low=np.asarray(sp.low).tolist(), high=np.asarray(sp.high).tolist())


def refgen_desc(rg):
    d = dict(kind=type(rg).__name__)
    if hasattr(rg, "_sub_generators"):
        d["sub_generators"] = [refgen_desc(s) for s in rg._sub_generators]
        return d
    for attr in ("_reference_state", "_limit_margin", "_episode_len_range", "_sigma_range", "_initial_range", "_reference_value"):
        if hasattr(rg, attr):
            d[attr.lstrip("_")] = jsonable(getattr(rg, attr))
    return d


def env_table():
    import gymnasium

    table = {}
    for env_id in sorted(k for k in gymnasium.envs.registration.registry.keys() if "-v0" in k and ("Cont-" in k or "Finite-" in k)):
        env = gem.make(env_id, visualization=NoViz())
        p = env.physical_system
        wrapped = p
        p = p.unwrapped
        load = p.mechanical_load
        conv = p.converter
        entry = dict(
            env_class=type(env.unwrapped).__name__,
            system_indices={k: jsonable(getattr(p, k)) for k in ("OMEGA_IDX", "TORQUE_IDX", "CURRENTS_IDX", "VOLTAGES_IDX", "U_SUP_IDX", "EPSILON_IDX")
                            if hasattr(p, k)},
            system_class=type(p).__name__,
            wrappers=[type(w).__name__ for w in _wrapper_chain(wrapped)],
            state_names=list(wrapped.state_names),
            limits=np.asarray(wrapped.limits).tolist(),
            nominal_state=np.asarray(wrapped.nominal_state).tolist(),
            state_low=wrapped.state_space.low.tolist(),
            state_high=wrapped.state_space.high.tolist(),
            action_space=space_desc(env.action_space),
            tau=float(p.tau),
            solver_class=type(p._ode_solver).__name__,
            motor_class=type(p.electrical_motor).__name__,
            motor_parameter={k: float(v) for k, v in p.electrical_motor.motor_parameter.items()},
            motor_limits=jsonable(p.electrical_motor.limits),
            motor_nominal=jsonable(p.electrical_motor.nominal_values),
            load_class=type(load).__name__,
            j_total=float(load.j_total),
            load_parameter=jsonable(getattr(load, "load_parameter", {})),
            omega_fixed=float(getattr(load, "omega_fixed", 0.0) or 0.0),
            supply_class=type(p.supply).__name__,
            u_sup=float(p.supply.u_nominal),
            converter_class=type(conv).__name__,
            sub_converters=[type(c).__name__ for c in getattr(conv, "_sub_converters", [])],
            interlocking_time=float(conv._interlocking_time),
            reference_generator=refgen_desc(env.reference_generator),
            reference_names=list(env.reference_generator.reference_names),
            reference_space=space_desc(env.reference_generator.reference_space),
            reward_weights=np.asarray(env.reward_function._reward_weights, dtype=float).tolist(),
            reward_power=np.asarray(env.reward_function._n, dtype=float).tolist(),
            reward_bias=float(env.reward_function._bias),
            violation_reward=float(env.reward_function._violation_reward),
            reward_range=[float(v) for v in env.reward_function.reward_range],
            constraints=[
                dict(
                    kind=type(c).__name__,
                    states=[str(s) for s in (getattr(c, "_states", None) or np.asarray(wrapped.state_names)[c._observed_states])],
                )
                for c in env.constraint_monitor.constraints
            ],
        )
        (s0, r0), _ = env.reset(seed=0)
        entry["reset_state"] = np.asarray(s0, dtype=float).tolist()
        table[env_id] = entry
    with open(os.path.join(HERE, "env_table.json"), "w") as f:
        json.dump(table, f, indent=1, sort_keys=True)
    print(f"env_table.json: {len(table)} ids")


def _wrapper_chain(ps_):
    chain = []
    while hasattr(ps_, "_physical_system") and ps_ is not ps_.unwrapped:
        chain.append(ps_)
        ps_ = ps_._physical_system
    return chain


def init_bounds():
    """Empirical range of the reference's random initial ODE states (random_init='uniform'), 3000 resets per env."""
    out = {}
    for env_id, load_iv in [("Cont-CC-PMSM-v0", None), ("Cont-SC-PMSM-v0", [[-50.0, 120.0]]), ("Cont-CC-EESM-v0", None),
                            ("Cont-CC-PermExDc-v0", None), ("Cont-SC-ExtExDc-v0", None), ("Cont-CC-SynRM-v0", None)]:
        env = gem.make(env_id, visualization=NoViz(), ode_solver=make_solver("euler"), motor=dict(motor_initializer=dict(random_init="uniform")),
                       load=dict(load_initializer=dict(random_init="uniform", interval=load_iv)))
        ys = []
        env.reset(seed=0)
        for _ in range(3000):
            env.reset()
            ys.append(ode_state(env))
        ys = np.array(ys)
        out[env_id] = dict(load_interval=load_iv, min=ys.min(axis=0).tolist(), max=ys.max(axis=0).tolist(), mean=ys.mean(axis=0).tolist())
        print(env_id, "ode min", np.round(ys.min(axis=0), 2), "max", np.round(ys.max(axis=0), 2))
    with open(os.path.join(HERE, "init_bounds.json"), "w") as f:
        json.dump(out, f, indent=1)


def init_bounds_induction():
    """Induction motors with random_init='uniform': the flux bounds are re-derived at every reset from a random field angle drawn from the GLOBAL
    numpy RNG, the speed and the previous reset's initial currents (squirrel_cage_induction_motor.py:146-157, induction_motor.py:250-285), so
    the device can only match the DISTRIBUTION: per ODE state min / max / mean / std over 20 000 resets (the first reset, which still sees
    the default currents, left out), plus mean and max of the flux magnitude and the share of exactly-zero flux."""
    out = {}
    np.random.seed(12345)
    for key, env_id, load_iv in [("sc_scim", "Cont-SC-SCIM-v0", None), ("cc_scim", "Cont-CC-SCIM-v0", None), ("sc_dfim", "Cont-SC-DFIM-v0", None),
                                 ("cc_dfim", "Cont-CC-DFIM-v0", None), ("sc_scim_randload", "Cont-SC-SCIM-v0", [[-50.0, 120.0]])]:
        load = dict(load_initializer=dict(random_init="uniform", interval=load_iv)) if load_iv else None
        env = gem.make(env_id, visualization=NoViz(), ode_solver=make_solver("euler"), motor=dict(motor_initializer=dict(random_init="uniform")), load=load)
        env.reset(seed=0)
        ys = []
        for _ in range(20000):
            env.reset()
            ys.append(ode_state(env))
        ys = np.array(ys)
        mag = np.hypot(ys[:, 3], ys[:, 4])
        out[key] = dict(env_id=env_id, load_interval=load_iv, min=ys.min(axis=0).tolist(), max=ys.max(axis=0).tolist(), mean=ys.mean(axis=0).tolist(),
                        std=ys.std(axis=0).tolist(), flux_mag_mean=float(mag.mean()), flux_mag_max=float(mag.max()), flux_zero_share=float((mag == 0).mean()))
        print(key, "ode min", np.round(ys.min(axis=0), 3), "max", np.round(ys.max(axis=0), 3), "flux |.| mean", round(float(mag.mean()), 4), "zero share", float((mag == 0).mean()))
    with open(os.path.join(HERE, "init_bounds_induction.json"), "w") as f:
        json.dump(out, f, indent=1)


def switched_stats():
    """SwitchedReferenceGenerator semantics as statistics (its numpy streams cannot be matched on a device): three constant
    sub-generators with probabilities (.5, .3, .2), super-episodes of integers(5, 12) steps, 200 000 steps of the reference; recorded:
    the histogram of run lengths of equal reference values, the value frequencies, the length of the first run after the reset."""
    from gym_electric_motor.reference_generators import ConstReferenceGenerator, SwitchedReferenceGenerator

    vals, probs, length = [0.3, -0.2, 0.6], [0.5, 0.3, 0.2], (5, 12)
    runs, first_runs, counts = [], [], {v: 0 for v in vals}
    for seed in range(40):
        env = gem.make("Cont-SC-PermExDc-v0", visualization=NoViz(), ode_solver=make_solver("euler"), constraints=(),
                       reference_generator=SwitchedReferenceGenerator([ConstReferenceGenerator("omega", v) for v in vals], p=probs,
                                                                      super_episode_length=length))
        (_, r0), _ = env.reset(seed=seed)
        seq = [float(r0[0])]
        for _ in range(5000):
            (_, rn), _, term, _, _ = env.step(np.zeros(1))
            assert not term
            seq.append(float(rn[0]))
        seq = np.array(seq)
        edges = np.nonzero(np.diff(seq) != 0)[0] + 1
        rl = np.diff(np.concatenate(([0], edges)))  # complete runs only (the last, unfinished one is dropped)
        first_runs.append(int(rl[0]))
        runs += rl[1:].tolist()
        for v in vals:
            counts[v] += int(np.isclose(seq, v).sum())
    hist = np.bincount(np.array(runs, dtype=int), minlength=80)[:80]
    out = dict(values=vals, p=probs, super_episode_length=list(length), run_length_hist=hist.tolist(), first_runs=first_runs,
               value_frequency=[counts[v] / sum(counts.values()) for v in vals], n_runs=len(runs))
    with open(os.path.join(HERE, "switched_stats.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("switched_stats: runs", len(runs), "min/max run", min(runs), max(runs), "first runs", sorted(set(first_runs)), "freq", out["value_frequency"])


def regen_ref_data():
    """Re-run the reference's own integration test recipe (tests/integration_tests/test_integration.py:18-87)."""
    sys.path.insert(0, os.path.join(REF_ROOT, "examples", "classic_controllers"))
    sys.path.insert(0, os.path.join(REF_ROOT, "src"))
    from classic_controllers import Controller  # reference example code, used here only as the oracle's driver
    from gym_electric_motor.reference_generators import SinusoidalReferenceGenerator

    motor_type, control_type, action_type = "PermExDc", "SC", "Cont"
    env_id = f"{action_type}-{control_type}-{motor_type}-v0"
    ref_generator = SinusoidalReferenceGenerator(
        amplitude_range=(1, 1), frequency_range=(5, 5), offset_range=(0, 0), episode_lengths=(10001, 10001)
    )
    env = gem.make(env_id, reference_generator=ref_generator, visualization=NoViz())
    controller = Controller.make(env)
    (state, reference), _ = env.reset(seed=1337)
    states, references, rewards, terms, truncs, actions = [], [], [], [], [], []
    for _ in range(2001):
        action = controller.control(state, reference)
        (state, reference), reward, terminated, truncated, _ = env.step(action)
        actions.append(np.atleast_1d(np.asarray(action, dtype=float)).copy())
        states.append(np.array(state))
        references.append(np.array(reference))
        rewards.append(reward)
        terms.append(terminated)
        truncs.append(truncated)
        if terminated:
            env.reset()
            controller.reset()
    states = np.array(states)
    gold = np.load(os.path.join(REF_ROOT, "tests", "integration_tests", "ref_data.npz"))
    n = min(len(states), len(gold["states"]))
    d = np.abs(states[:n] - gold["states"][:n]).max()
    assert np.allclose(states[:n], gold["states"][:n]), d
    assert np.allclose(np.array(references)[:n], gold["references"][:n])
    assert np.allclose(np.array(rewards)[:n], gold["rewards"][:n])
    print(f"ref_data.npz reproduced by the reference run here: max|dstates|={d:.3e}")
    meta = describe(env, dict(name="ref_data_regen", env_id=env_id, solver="dopri5", seed=1337))
    np.savez_compressed(
        os.path.join(HERE, "ref_data_regen.npz"),
        actions=np.array(actions),
        states=states,
        references=np.array(references),
        rewards=np.array(rewards),
        terminated=np.array(terms, dtype=np.uint8),
        meta=json.dumps(meta),
    )


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--skip-table", action="store_true")
    args = ap.parse_args()
    for case in CASES:
        if args.only in ("ref_data", "init_bounds", "init_bounds_induction", "switched_stats", "env_table") or (args.only and args.only not in case["name"]):
            continue
        record(case)
    if not args.only or args.only == "init_bounds_induction":
        init_bounds_induction()
    if not args.only or args.only == "init_bounds":
        init_bounds()
    if not args.only or args.only == "switched_stats":
        switched_stats()
    if args.only == "env_table":
        env_table()
    if not args.only or args.only == "ref_data":
        if not args.skip_table and not args.only:
            env_table()
        regen_ref_data()

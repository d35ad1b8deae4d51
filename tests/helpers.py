"""Shared test helpers: golden-fixture loading and golden-meta -> gemb200_config translation.

The translation here is deliberately independent of the product's own spec compiler
(gym_electric_motor_b200/spec.py): it takes limits / weights verbatim from the values the REFERENCE reported
when the golden was recorded, so oracle-vs-golden tests do not depend on the product's host logic.
"""
import glob
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from gym_electric_motor_b200 import _cabi as K  # noqa: E402

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")

MOTOR_KIND = {
    "DcPermanentlyExcitedMotor": K.MOTOR_PERMEX_DC,
    "DcSeriesMotor": K.MOTOR_SERIES_DC,
    "DcShuntMotor": K.MOTOR_SHUNT_DC,
    "DcExternallyExcitedMotor": K.MOTOR_EXTEX_DC,
    "PermanentMagnetSynchronousMotor": K.MOTOR_PMSM,
    "SynchronousReluctanceMotor": K.MOTOR_SYNRM,
    "ExternallyExcitedSynchronousMotor": K.MOTOR_EESM,
    "SquirrelCageInductionMotor": K.MOTOR_SCIM,
    "DoublyFedInductionMotor": K.MOTOR_DFIM,
}
MP_SLOT = dict(
    p=K.MP_P, r_s=K.MP_R_S, l_d=K.MP_L_D, l_q=K.MP_L_Q, psi_p=K.MP_PSI_P, j_rotor=K.MP_J_ROTOR, r_a=K.MP_R_A,
    l_a=K.MP_L_A, psi_e=K.MP_PSI_E, r_e=K.MP_R_E, l_e=K.MP_L_E, l_e_prime=K.MP_L_E_PRIME, l_m=K.MP_L_M, k=K.MP_K,
    l_sigs=K.MP_L_SIGS, l_sigr=K.MP_L_SIGR, r_r=K.MP_R_E,
)
CONV = {
    "ContOneQuadrantConverter": (0, [K.CONV_1QC]),
    "ContTwoQuadrantConverter": (0, [K.CONV_2QC]),
    "ContFourQuadrantConverter": (0, [K.CONV_4QC]),
    "ContB6BridgeConverter": (0, [K.CONV_B6]),
    "FiniteOneQuadrantConverter": (1, [K.CONV_1QC]),
    "FiniteTwoQuadrantConverter": (1, [K.CONV_2QC]),
    "FiniteFourQuadrantConverter": (1, [K.CONV_4QC]),
    "FiniteB6BridgeConverter": (1, [K.CONV_B6]),
}
N_MOTOR_ODE = {K.MOTOR_PERMEX_DC: 1, K.MOTOR_SERIES_DC: 1, K.MOTOR_SHUNT_DC: 2, K.MOTOR_EXTEX_DC: 2, K.MOTOR_PMSM: 3,
               K.MOTOR_SYNRM: 3, K.MOTOR_EESM: 4, K.MOTOR_SCIM: 5, K.MOTOR_DFIM: 5}


def golden_names():
    return sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if "ref_data" not in p)


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files if k != "meta"}
    d["meta"] = json.loads(str(z["meta"]))
    return d


def solver_from_name(name):
    """'euler', 'euler3', 'rk4', 'rk4x2', 'dopri5' -> (kind, nsteps); dopri5 is oracle-only (kind 100)."""
    if name == "dopri5":
        return 100, 1
    if name.startswith("euler"):
        return K.SOLVER_EULER, int(name[5:] or 1)
    if name.startswith("rk4x"):
        return K.SOLVER_RK4, int(name[4:])
    if name == "rk4":
        return K.SOLVER_RK4, 1
    raise ValueError(name)


def len_steps_hint(meta):
    return int(meta["case"].get("steps", 2000))


def config_from_meta(meta, n_envs=1, solver=None, ref_kind=K.REF_EXTERNAL, dtype=K.F64, layout=K.LAYOUT_AOS,
                     autoreset=K.AUTORESET_NONE, seed=0, reset_ode=None):
    cfg = K.new_config()
    cfg.n_envs = n_envs
    cfg.dtype, cfg.layout, cfg.autoreset = dtype, layout, autoreset
    mk = MOTOR_KIND[meta["motor_class"]]
    cfg.motor_kind = mk
    cc = meta["converter_class"]
    if cc in CONV:
        cfg.finite, kinds = CONV[cc]
    elif cc in ("ContMultiConverter", "FiniteMultiConverter"):
        cfg.finite = int(cc.startswith("Finite"))
        sub = {"ContFourQuadrantConverter": K.CONV_4QC, "ContTwoQuadrantConverter": K.CONV_2QC, "ContOneQuadrantConverter": K.CONV_1QC,
               "FiniteFourQuadrantConverter": K.CONV_4QC, "FiniteTwoQuadrantConverter": K.CONV_2QC, "FiniteOneQuadrantConverter": K.CONV_1QC,
               "ContB6BridgeConverter": K.CONV_B6, "FiniteB6BridgeConverter": K.CONV_B6}
        multi = meta["case"].get("multi")
        if multi:
            kinds = [sub[name] for name, _ in multi]
        else:
            kinds = [K.CONV_B6, K.CONV_4QC] if mk == K.MOTOR_EESM else ([K.CONV_B6, K.CONV_B6] if mk == K.MOTOR_DFIM else [K.CONV_4QC, K.CONV_4QC])
    else:
        raise ValueError(cc)
    for i, kd in enumerate(kinds):
        cfg.converter_kind[i] = kd
    cfg.load_kind = {"ConstantSpeedLoad": K.LOAD_CONST_SPEED, "ExternalSpeedLoad": K.LOAD_EXT_SPEED}.get(meta["load_class"], K.LOAD_POLY_STATIC)
    kind, nsteps = solver_from_name(solver or meta["case"]["solver"])
    cfg.solver_kind, cfg.solver_nsteps = kind, nsteps
    cfg.tau = meta["tau"]
    if cfg.load_kind == K.LOAD_EXT_SPEED:  # tabulated speed profile: f(j tau / (2 nsteps) + tau_load), make_golden.py:sin_profile
        es = meta["ext_speed"]
        per = 2 * nsteps
        j = np.arange(per * (len_steps_hint(meta) + 8) + 2 * per + 1)
        t = j * (meta["tau"] / per) + es["tau"]
        tab = np.ascontiguousarray(es["o"] + es["a"] * np.sin(2 * np.pi * es["f"] * t))
        cfg.load_param[K.LP_TAU_LOAD] = es["tau"]
        cfg.ext_speed_table, cfg.ext_speed_len = tab.ctypes.data, len(tab)
        cfg._keepalive = tab
    cfg.interlocking_time = meta["interlocking_time"]
    ils = meta.get("interlocking_times") or []
    if len(ils) == 2 and ils[0] != ils[1]:  # multi converter whose sub-converters differ: one time per converter slot
        cfg.interlocking_time, cfg.interlocking_time1 = ils[0], ils[1]
    cfg.u_sup = meta["u_sup"]
    if meta.get("supply_class") == "AC1PhaseSupply":
        cfg.supply_kind = K.SUPPLY_AC1
        cfg.supply_param[0], cfg.supply_param[1], cfg.supply_param[2] = meta["supply_parameter"]["f"], meta["supply_parameter"]["phase"], 1.0
    if meta.get("supply_class") == "RCVoltageSupply":
        cfg.supply_kind = K.SUPPLY_RC
        cfg.supply_param[0], cfg.supply_param[1] = meta["supply_parameter"]["R"], meta["supply_parameter"]["C"]
    for k, v in meta["motor_parameter"].items():
        if k in MP_SLOT:
            cfg.motor_param[MP_SLOT[k]] = v
    lp = meta.get("load_parameter") or {}
    cfg.load_param[K.LP_A] = lp.get("a", 0.0)
    cfg.load_param[K.LP_B] = lp.get("b", 0.0)
    cfg.load_param[K.LP_C] = lp.get("c", 0.0)
    # j_total = j_load + j_rotor (mechanical_load.py:188-193); golden meta stores j_total
    cfg.load_param[K.LP_J_LOAD] = meta["j_total"] - meta["motor_parameter"]["j_rotor"]
    cfg.load_param[K.LP_TAU_DECAY] = 1e-3
    n_state = len(meta["state_names"])
    # cfg.limits normalises the inner system's own vector; weights / lengths refer to the final (wrapped) vector
    for i, v in enumerate(meta.get("base_limits", meta["limits"])):
        cfg.limits[i] = v
    for i in range(n_state):
        cfg.reward_weight[i] = meta["reward_weights"][i]
        cfg.reward_power[i] = meta["reward_power"][i]
        cfg.state_length[i] = meta["state_length"][i]
    cfg.reward_bias = meta["reward_bias"]
    cfg.violation_reward = meta["violation_reward"]
    n_ode = 1 + N_MOTOR_ODE[mk]
    if reset_ode is not None:
        for i in range(n_ode):
            cfg.init_ode[i] = float(reset_ode[i])
    names = meta["state_names"]
    cfg.n_constraints = len(meta["constraints"])
    for ci, con in enumerate(meta["constraints"]):
        cfg.constraint_kind[ci] = K.CONSTRAINT_SQUARED if con["kind"] == "SquaredConstraint" else K.CONSTRAINT_LIMIT
        m = 0
        for s in con["states"]:
            m |= 1 << names.index(s)
        cfg.constraint_mask[ci] = m
    ref_names = meta["reference_names"]
    cfg.n_ref = len(ref_names)
    for r, rn in enumerate(ref_names):
        cfg.ref_kind[r] = ref_kind
        cfg.ref_state[r] = names.index(rn)
    cfg.seed = seed
    # physical-system wrappers recorded with the golden (list order = reference order: later entries wrap earlier ones)
    dq, dead, outer, adv = 0, 0, 0, 0.0
    cur_names = list(meta.get("base_state_names", names))  # state vector as seen by the next wrapper
    cur_limits = list(meta.get("base_limits", meta["limits"]))
    nops = 0
    for kind, arg in meta["case"].get("wrappers", []) or []:
        if kind == "DeadTime":
            dead, outer = int(arg), (1 if dq else 0)
        elif kind == "CosSin":  # cos_sin_processor.py:39-58
            idx, rm = cur_names.index(arg[0]), int(arg[1])
            cfg.sop_kind[nops] = K.SOP_COS_SIN
            cfg.sop_idx[nops][0], cfg.sop_idx[nops][1] = idx, rm
            if rm:
                del cur_names[idx], cur_limits[idx]
            cur_names += [f"cos({arg[0]})", f"sin({arg[0]})"]
            cur_limits += [1.0, 1.0]
            nops += 1
        elif kind == "FluxObserver":  # flux_observer.py:56-79
            mp = meta["motor_parameter"]
            l_r = mp["l_m"] + mp["l_sigr"]
            psi_limit = mp["l_m"] * cur_limits[cur_names.index("i_sd")]
            idx = [cur_names.index(n) for n in ("i_sa", "i_sb", "i_sc", "omega")]
            cfg.sop_kind[nops] = K.SOP_FLUX_OBSERVER
            for q, v in enumerate(idx):
                cfg.sop_idx[nops][q] = v
            for q, v in enumerate([mp["r_r"] * mp["l_m"] / l_r, mp["r_r"] / l_r, mp["p"], psi_limit] + [cur_limits[j] for j in idx]):
                cfg.sop_param[nops][q] = v
            cur_names += ["psi_abs", "psi_angle"]
            cur_limits += [psi_limit, np.pi]
            nops += 1
        else:
            dq, adv = (3 if arg == "DFIM" else (2 if arg == "SCIM" else 1)), 0.5 + dead
    cfg.n_state_ops = nops
    assert cur_names == names or not nops, (cur_names, names)
    cfg.action_dq, cfg.dead_time_steps, cfg.dead_time_outer, cfg.angle_advance = dq, dead, outer, adv
    return cfg


def replay_golden(sim, g, inject_refs=True):
    """Drive `sim` (Oracle or device env with the same reset/step/set_reference API, N=1) through a golden
    trajectory: same actions, reference injected before every step, reset on termination exactly where the
    reference harness did (tests/golden/make_golden.py:record).  Returns dict of arrays shaped like the golden."""
    K_ = len(g["actions"])
    ref_idx = [g["meta"]["state_names"].index(n) for n in g["meta"]["reference_names"]]
    obs0, _ = sim.reset()
    out = dict(states=np.zeros_like(g["states"]), rewards=np.zeros(K_), terminated=np.zeros(K_, dtype=np.uint8),
               reset_state=np.asarray(obs0, dtype=np.float64)[0])
    for k in range(K_):
        if inject_refs and ref_idx:
            sim.set_reference(g["refs_used"][k][ref_idx][None, :])
        obs, ref, rew, term = sim.step(g["actions"][k][None, ...] if g["actions"].ndim > 1 else g["actions"][k : k + 1])
        out["states"][k] = np.asarray(obs, dtype=np.float64)[0]
        out["rewards"][k] = float(np.asarray(rew)[0])
        out["terminated"][k] = int(np.asarray(term)[0])
        if g["terminated"][k]:
            sim.reset()
    return out


def switched_config(n, kinds_cfg, p, length, name="permex_sc_euler3", seed=3, dtype=K.F64):
    """config with ONE switched reference slot on the golden's referenced state; kinds_cfg = list of dicts written into the extra
    parameter entries 1.. (entry 0 is the output slot itself)."""
    g = load_golden(name)
    cfg = config_from_meta(g["meta"], n_envs=n, reset_ode=g["reset_ode"], solver="rk4", ref_kind=K.REF_WIENER, seed=seed, dtype=dtype,
                           autoreset=K.AUTORESET_SAME_STEP)
    assert cfg.n_ref == 1
    cfg.n_constraints = 0
    cfg.ref_sw_count[0], cfg.ref_sw_first[0] = len(kinds_cfg), 1
    cfg.ref_sw_len_lo[0], cfg.ref_sw_len_hi[0] = length
    acc = 0.0
    for j, kc in enumerate(kinds_cfg):
        e = 1 + j
        cfg.ref_state[e] = cfg.ref_state[0]
        cfg.ref_kind[e] = kc["kind"]
        cfg.ref_value[e] = kc.get("value", 0.0)
        cfg.ref_margin_lo[e], cfg.ref_margin_hi[e] = kc.get("margin", (-0.8, 0.8))
        cfg.ref_init_lo[e], cfg.ref_init_hi[e] = kc.get("margin", (-0.8, 0.8))
        cfg.ref_sigma_lo[e], cfg.ref_sigma_hi[e] = kc.get("sigma", (1e-3, 1e-2))
        cfg.ref_len_lo[e], cfg.ref_len_hi[e] = kc.get("length", (8, 30))
        cfg.ref_amp_lo[e], cfg.ref_amp_hi[e] = kc.get("amp", (0.1, 0.4))
        cfg.ref_freq_lo[e], cfg.ref_freq_hi[e] = kc.get("freq", (50.0, 400.0))
        cfg.ref_off_lo[e], cfg.ref_off_hi[e] = kc.get("off", (-0.3, 0.3))
        acc += p[j]
        cfg.ref_sw_cdf[e] = acc if j < len(kinds_cfg) - 1 else 1.0
    return cfg


def golden_reset_state(g):
    """Reset observation of a golden.  Reference quirk: CosSinProcessor(remove_angle=True).reset() returns the vector WITH the
    angle it removes in simulate() (cos_sin_processor.py:60-63 vs :65-70); a batched tensor has one width, so the device path
    and the oracle remove it at reset too — compare against the golden with that column deleted."""
    rs = np.asarray(g["reset_state"], dtype=np.float64)
    names = list(g["meta"].get("base_state_names", g["meta"]["state_names"]))
    for kind, arg in g["meta"]["case"].get("wrappers", []) or []:
        if kind == "CosSin":
            if int(arg[1]):  # env.reset() then cuts the un-shortened vector with its state filter: [..., angle, ..., cos] — sin is lost
                a = names.index(arg[0])
                rs = np.delete(np.concatenate((rs, [np.sin(np.pi * rs[a])])), a)
            if int(arg[1]):
                names.remove(arg[0])
            names += [f"cos({arg[0]})", f"sin({arg[0]})"]
        elif kind == "FluxObserver":
            names += ["psi_abs", "psi_angle"]
    return rs


def col_rel_err(a, b):
    """max over columns of max|a-b| / max(|b|) — the 'column-relative' error of SURVEY.md §7."""
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    scale = np.maximum(np.abs(b).max(axis=0), 1e-12)
    return float((np.abs(a - b).max(axis=0) / scale).max())

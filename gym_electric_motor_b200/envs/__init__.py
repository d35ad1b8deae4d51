"""Registry of the reference's environment ids and their default components.

The reference registers 54 ids `{Cont|Finite}-{CC|TC|SC}-{motor}-v0` (src/gym_electric_motor/__init__.py:44-283), each a
thin class that only picks default components (envs/**).  Here the defaults are DATA (rules + the per-env exceptions
listed in SURVEY.md Appendix A); tests/test_host_envs.py diffs every derived quantity against
tests/golden/env_table.json, which was dumped from the running reference.

`make(env_id, num_envs=N, device=0, dtype='float32', **reference_kwargs)` accepts the reference's env kwargs
(supply, converter, motor, load, ode_solver, reward_function, reference_generator, constraints, tau, state_filter,
callbacks, …) with the same `env-arg` semantics: None -> default, instance -> used as is, dict -> default class with
updated kwargs (utils.initialize).
"""
from .. import physical_systems as ps
from ..constraints import SquaredConstraint
from ..core import ElectricMotorEnvironment, ElectricMotorVisualization
from ..reference_generators import MultipleReferenceGenerator, ReferenceGenerator, WienerProcessReferenceGenerator
from ..reward_functions import RewardFunction, WeightedSumOfErrors
from ..utils import initialize
from .motors import ActionType, ControlType, Motor, MotorType  # noqa: F401

MOTORS = ["PermExDc", "SeriesDc", "ShuntDc", "ExtExDc", "PMSM", "SynRM", "EESM", "SCIM", "DFIM"]
_DC = ("PermExDc", "SeriesDc", "ShuntDc", "ExtExDc")

_MOTOR_CLASS = dict(PermExDc=ps.DcPermanentlyExcitedMotor, SeriesDc=ps.DcSeriesMotor, ShuntDc=ps.DcShuntMotor, ExtExDc=ps.DcExternallyExcitedMotor,
                    PMSM=ps.PermanentMagnetSynchronousMotor, SynRM=ps.SynchronousReluctanceMotor, EESM=ps.ExternallyExcitedSynchronousMotor,
                    SCIM=ps.SquirrelCageInductionMotor, DFIM=ps.DoublyFedInductionMotor)
_SYSTEM_CLASS = dict(PermExDc=ps.DcMotorSystem, SeriesDc=ps.DcMotorSystem, ShuntDc=ps.DcMotorSystem, ExtExDc=ps.DcMotorSystem,
                     PMSM=ps.SynchronousMotorSystem, SynRM=ps.SynchronousMotorSystem, EESM=ps.ExternallyExcitedSynchronousMotorSystem,
                     SCIM=ps.SquirrelCageInductionMotorSystem, DFIM=ps.DoublyFedInductionMotorSystem)
_CC_STATES = dict(PermExDc=("i",), SeriesDc=("i",), ShuntDc=("i_a",), ExtExDc=("i_a", "i_e"), PMSM=("i_sd", "i_sq"), SynRM=("i_sd", "i_sq"),
                  EESM=("i_sd", "i_sq", "i_e"), SCIM=("i_sd", "i_sq"), DFIM=("i_sd", "i_sq"))
# sigma_range of the omega Wiener reference in the SC envs: (Cont, Finite)
_SC_SIGMA = dict(PermExDc=((1e-3, 5e-2), (1e-3, 5e-3)), SeriesDc=((1e-3, 2e-2), (1e-3, 5e-3)), ShuntDc=((1e-3, 3e-2), (1e-3, 5e-3)),
                 ExtExDc=((1e-3, 1e-1), (1e-3, 1e-1)), PMSM=((1e-3, 1e-1), (1e-3, 1e-1)), SynRM=((1e-3, 1e-2), (1e-3, 1e-2)),
                 EESM=((1e-3, 1e-1), (1e-3, 1e-1)), SCIM=((1e-3, 1e-2), (1e-3, 1e-2)), DFIM=((1e-3, 1e-2), (1e-3, 1e-2)))


def env_ids():
    return [f"{a}-{c}-{m}-v0" for a in ("Cont", "Finite") for c in ("CC", "TC", "SC") for m in MOTORS]


def parse_env_id(env_id):
    try:
        a, c, m, v = env_id.split("-")
    except ValueError:
        raise KeyError(f"unknown environment id {env_id!r}") from None
    if a not in ("Cont", "Finite") or c not in ("CC", "TC", "SC") or m not in MOTORS or v != "v0":
        raise KeyError(f"unknown environment id {env_id!r}")
    return a, c, m


def _default_converter(a, m):
    cont = a == "Cont"
    qc4 = ps.ContFourQuadrantConverter if cont else ps.FiniteFourQuadrantConverter
    b6 = ps.ContB6BridgeConverter if cont else ps.FiniteB6BridgeConverter
    multi = ps.ContMultiConverter if cont else ps.FiniteMultiConverter
    if m in ("PermExDc", "SeriesDc", "ShuntDc"):
        return qc4, dict()
    # the reference's envs pass INSTANCES here (e.g. envs/gym_eesm/cont_cc_eesm_env.py:155-158), so converter=dict(interlocking_time=..)
    # reaches only the multi converter's unused copy, not the sub-converters; kept
    if m == "ExtExDc":
        return multi, dict(subconverters=(qc4(), qc4()))
    if m == "EESM":
        return multi, dict(subconverters=(b6(), qc4()))
    if m == "DFIM":
        return multi, dict(subconverters=(b6(), b6()))
    return b6, dict()


def _default_u_sup(a, c, m):
    if m in _DC:
        return 420.0 if (m == "SeriesDc" and a == "Finite" and c in ("CC", "TC")) else 60.0
    return 300.0 if (a == "Cont" and c == "CC" and m in ("PMSM", "EESM")) else 420.0


def _default_load(a, c, m):
    if c in ("CC", "TC"):
        return ps.ConstantSpeedLoad, dict(omega_fixed=230.0 if (a, c, m) == ("Cont", "TC", "ShuntDc") else 100.0)
    if m == "PermExDc":
        lp = dict(a=0.0, b=0.0, c=0.0, j_load=1e-4 if a == "Cont" else 1e-3)
    elif m == "ExtExDc":
        lp = dict(a=0.0, b=0.0, c=0.0, j_load=1e-4)
    elif m == "ShuntDc":
        lp = dict(a=0.05, b=0.01, c=0.0, j_load=1e-4)
    elif m == "SeriesDc":
        lp = dict(a=0.01, b=0.05, c=0.0, j_load=1e-4) if a == "Cont" else dict(a=0.15, b=0.05, c=0.0, j_load=1e-4)
    elif (a, m) == ("Finite", "EESM"):
        lp = dict(a=0.0, b=0.0, c=0.0, j_load=1e-5)
    else:
        lp = dict(a=0.01, b=0.01, c=0.0, j_load=1e-5)
    return ps.PolynomialStaticLoad, dict(load_parameter=lp)


def _default_reference(a, c, m):
    permex_sigma = dict(sigma_range=(1e-2, 1e-1)) if m == "PermExDc" else {}
    if c == "CC":
        states = _CC_STATES[m]
        subs = []
        for s in states:
            kw = dict(reference_state=s, **permex_sigma)
            if (a, m, s) == ("Cont", "EESM", "i_e"):
                kw["limit_margin"] = (0, 1)
            subs.append(WienerProcessReferenceGenerator(**kw))
        if len(subs) == 1:
            return WienerProcessReferenceGenerator, dict(reference_state=states[0], **permex_sigma)
        return MultipleReferenceGenerator, dict(sub_generators=tuple(subs))
    if c == "TC":
        kw = dict(reference_state="torque", **permex_sigma)
        if (a, m) == ("Cont", "ShuntDc"):
            kw["limit_margin"] = (0, 0.8)
        return WienerProcessReferenceGenerator, kw
    return WienerProcessReferenceGenerator, dict(reference_state="omega", sigma_range=_SC_SIGMA[m][0 if a == "Cont" else 1])


def _default_reward_weights(c, m):
    """the env classes name their reward weights explicitly — the controlled currents in equal shares, omega or torque (e.g.
    envs/gym_pmsm/cont_cc_pmsm_env.py:172, gym_eesm/cont_cc_eesm_env.py:183, gym_pmsm/cont_sc_pmsm_env.py:169) — so a user-supplied
    reference generator on another state does NOT move the default reward with it"""
    if c == "SC":
        return dict(omega=1.0)
    if c == "TC":
        return dict(torque=1.0)
    return {name: 1.0 / len(_CC_STATES[m]) for name in _CC_STATES[m]}


def _default_constraints(m):
    if m in ("PermExDc", "SeriesDc"):
        return ("i",)
    if m in ("ShuntDc", "ExtExDc"):
        return ("i_a", "i_e")
    if m == "EESM":
        return (SquaredConstraint(("i_sq", "i_sd")), "i_e")
    return (SquaredConstraint(("i_sq", "i_sd")),)


_NOT_SET = object()

# The reference has one class per id (envs/**: `ContCurrentControlPermanentMagnetSynchronousMotorEnv`, ...) and agents test
# `type(env) in (envs.ContSpeedControlDcExternallyExcitedMotorEnv, ...)`; here the defaults are data, so the 54 names are thin
# subclasses generated from the same table and `make` instantiates the one that belongs to the id.
_ACTION_WORD = dict(Cont="Cont", Finite="Finite")
_CONTROL_WORD = dict(CC="CurrentControl", TC="TorqueControl", SC="SpeedControl")
ENV_CLASSES = {}
for _a in _ACTION_WORD:
    for _c in _CONTROL_WORD:
        for _m in MOTORS:
            _cls_name = f"{_ACTION_WORD[_a]}{_CONTROL_WORD[_c]}{_MOTOR_CLASS[_m].__name__}Env"
            ENV_CLASSES[f"{_a}-{_c}-{_m}-v0"] = type(_cls_name, (ElectricMotorEnvironment,), {"__doc__": f"`{_a}-{_c}-{_m}-v0` (defaults: envs table)"})
            globals()[_cls_name] = ENV_CLASSES[f"{_a}-{_c}-{_m}-v0"]


def make(env_id, supply=None, converter=None, motor=None, load=None, ode_solver=None, reward_function=None, reference_generator=None,
         visualization=None, state_filter=None, callbacks=(), constraints=_NOT_SET, calc_jacobian=True, tau=None,
         physical_system_wrappers=(), num_envs=None, device=0, dtype="float32", layout="aos", autoreset=None, seed=None,
         env_index_offset=0, **kwargs):
    """`gem.make` for the device path (reference core.py:291-292 -> env constructors, e.g.
    envs/gym_pmsm/cont_cc_pmsm_env.py:95-190)."""
    a, c, m = parse_env_id(env_id)
    tau = (1e-4 if a == "Cont" else 1e-5) if tau is None else tau
    conv_cls, conv_args = _default_converter(a, m)
    load_cls, load_args = _default_load(a, c, m)
    ref_cls, ref_args = _default_reference(a, c, m)
    n = 1 if num_envs is None else int(num_envs)
    physical_system = _SYSTEM_CLASS[m](
        supply=initialize(ps.VoltageSupply, supply, ps.IdealVoltageSupply, dict(u_nominal=_default_u_sup(a, c, m))),
        converter=initialize(ps.PowerElectronicConverter, converter, conv_cls, conv_args),
        motor=initialize(ps.ElectricMotor, motor, _MOTOR_CLASS[m], dict()),
        load=initialize(ps.MechanicalLoad, load, load_cls, load_args),
        ode_solver=initialize(ps.OdeSolver, ode_solver, ps.ScipyOdeSolver, dict()),
        calc_jacobian=calc_jacobian, tau=tau, num_envs=n, device=device, dtype=dtype, layout=layout, env_index_offset=env_index_offset,
    )
    reference_generator = initialize(ReferenceGenerator, reference_generator, ref_cls, ref_args)
    reward_function = initialize(RewardFunction, reward_function, WeightedSumOfErrors, dict(reward_weights=_default_reward_weights(c, m)))
    if constraints is _NOT_SET:
        constraints = _default_constraints(m)
    if visualization is not None and not isinstance(visualization, (ElectricMotorVisualization, list, tuple)):
        visualization = None  # dict/str specs of the matplotlib dashboard: plotting is out of scope, ignored
    env = ENV_CLASSES[env_id](
        physical_system=physical_system, reference_generator=reference_generator, reward_function=reward_function,
        constraints=constraints, visualization=visualization or (), state_filter=state_filter, callbacks=callbacks,
        physical_system_wrappers=physical_system_wrappers, num_envs=num_envs, autoreset=autoreset, seed=seed, **kwargs)
    env.env_id = env_id
    return env

"""Enum helper for composing environment ids, with the names of the reference's `envs/motors.py` (an agent-side convenience:
`Motor(MotorType.PermanentMagnetSynchronousMotor, ControlType.TorqueControl, ActionType.Continuous).env_id()` -> 'Cont-TC-PMSM-v0').
Built from one table: member name -> (env-id tag, plotted state names)."""
from dataclasses import dataclass
from enum import Enum

_ABC = ["a", "b", "c"]
_DQ = ["d", "q"]


def _three_phase(prefix_i, prefix_u, order):
    """state names of a stator (or rotor) circuit in the order the reference lists them for plotting"""
    names = []
    for quantity in (prefix_i, prefix_u):
        for group in order:
            names += [f"{quantity}{k}" for k in group]
    return names


_MOTOR_TABLE = {
    "PermanentlyExcitedDcMotor": ("PermExDc", ["omega", "torque", "i", "u"]),
    "ExternallyExcitedDcMotor": ("ExtExDc", ["omega", "torque", "i_a", "i_e", "u_a", "u_e"]),
    "SeriesDc": ("SeriesDc", ["omega", "torque", "i", "u"]),
    "ShuntDc": ("ShuntDc", ["omega", "torque", "i_a", "i_e", "u"]),
    "ExternallyExcitedSynchronousMotor": ("EESM", ["omega", "torque", "i_sd", "i_sq", "i_a", "i_b", "i_c", "i_e", "u_sd", "u_sq", "u_a", "u_b", "u_c", "u_e"]),
    "DoublyFedInductionMotor": ("DFIM", ["omega", "torque"] + _three_phase("i_s", "u_s", (_ABC, _DQ)) + _three_phase("i_r", "u_r", (_ABC, _DQ)) + ["epsilon"]),
    "SquirrelCageInductionMotor": ("SCIM", ["omega", "torque"] + _three_phase("i_s", "u_s", (_ABC, _DQ)) + ["epsilon"]),
    "PermanentMagnetSynchronousMotor": ("PMSM", ["omega", "torque", "i_sd", "i_sq", "i_a", "i_b", "i_c", "u_sd", "u_sq", "u_a", "u_b", "u_c"]),
    "SynchronousReluctanceMotor": ("SynRM", ["omega", "torque", "i_sd", "i_sq", "i_a", "i_b", "i_c", "u_sd", "u_sq", "u_a", "u_b", "u_c"]),
}

MotorType = Enum("MotorType", list(_MOTOR_TABLE))
for _name, (_tag, _states) in _MOTOR_TABLE.items():
    MotorType[_name].env_id_tag = _tag
    MotorType[_name].states = list(_states)

ControlType = Enum("ControlType", ["SpeedControl", "TorqueControl", "CurrentControl"])
for _member, _tag in zip(ControlType, ("SC", "TC", "CC")):
    _member.env_id_tag = _tag

ActionType = Enum("ActionType", ["Continuous", "Finite"])
for _member, _tag in zip(ActionType, ("Cont", "Finite")):
    _member.env_id_tag = _tag


@dataclass
class Motor:
    motor_type: MotorType
    control_type: ControlType
    action_type: ActionType

    def env_id(self) -> str:
        return "-".join(part.env_id_tag for part in (self.action_type, self.control_type, self.motor_type)) + "-v0"

    def states(self) -> list:
        return self.motor_type.states

"""Host-side helpers with the semantics of the reference's utils.py (env-arg resolution, state arrays, parameter dicts)."""
import numpy as np


def initialize(base_class, arg, default_class, default_args):
    """env-arg resolver: None -> default, instance -> as is, dict -> default class with updated kwargs,
    str/type -> exception (reference utils.py:5-16)."""
    if arg is None:
        return default_class(**default_args)
    if isinstance(arg, type):
        raise Exception("Need initialization value")
    if isinstance(arg, base_class):
        return arg
    if type(arg) is str:
        raise Exception("Deprecated in version 3.0.0")
    if type(arg) is dict:
        args = dict(default_args)
        args.update(arg)
        return default_class(**args)
    raise TypeError(f"cannot build a {base_class.__name__} from {type(arg).__name__}: only the built-in component classes of "
                    "gym_electric_motor_b200 can run on the device (see INTEGRATION.md)")


def set_state_array(input_values, state_names):
    """dict / list / ndarray / scalar -> array over state_names (reference utils.py:40-70)."""
    if isinstance(input_values, dict):
        low = {str(k).lower(): v for k, v in input_values.items()}
        assert all(k in state_names for k in low), f"A state name in {low.keys()} is invalid."
        return np.array([low.get(n, 0.0) for n in state_names], dtype=float)
    if isinstance(input_values, (np.ndarray, list)):
        assert len(input_values) == len(state_names)
        return np.asarray(input_values, dtype=float)
    if isinstance(input_values, (float, int)):
        return input_values * np.ones(len(state_names), dtype=float)
    raise Exception("Incorrect type for the input values.")


def update_parameter_dict(source_dict, update_dict, copy=True):
    """dict.update that raises KeyError on unknown keys (reference utils.py:73-94)."""
    for key in update_dict.keys():
        if key not in source_dict:
            raise KeyError(f'Cannot update_dict the source_dict. The key "{key}" is not available.')
    new_dict = source_dict.copy() if copy else source_dict
    new_dict.update(update_dict)
    return new_dict


def state_dict_to_state_array(state_dict, state_array, state_names):
    """write the entries of a {state name: value} dict into `state_array` in place; names are case-insensitive, unknown names assert
    (reference utils.py:19-37)."""
    low = {str(k).lower(): v for k, v in state_dict.items()}
    assert all(k in state_names for k in low), f"A state name in {list(low)} is invalid."
    for ind, key in enumerate(state_names):
        if key in low:
            state_array[ind] = low[key]

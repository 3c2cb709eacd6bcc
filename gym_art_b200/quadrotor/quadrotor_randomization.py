"""Model samplers and parameter sanity limits (host side).

API of the reference's gym_art/quadrotor/quadrotor_randomization.py that the env-step path needs:
check_quad_param_limits (:16-46), the base samplers Crazyflie / DefaultQuad / MediumQuad (:329-339) through
which QuadrotorEnv names a model (quadrotor.py:738), and ConstValueSampler (:371-377).
Per-env dynamics randomisation (RandomQuad, Relative/AbsoluteSampler) is SURVEY.md section 8 row f2 (next).
"""
import copy

import numpy as np

from . import quad_models


def _update_existing(dst, upd):
    for k, v in upd.items():
        if isinstance(dst[k], dict):
            _update_existing(dst[k], v)
        else:
            dst[k] = v


def check_quad_param_limits(params, params_init=None):
    g, d, m = params["geom"], params["damp"], params["motor"]
    for part in ("body", "payload", "arms", "motors", "propellers"):      # lengths and masses are non-negative
        g[part] = {k: np.clip(v, 0., None) for k, v in g[part].items()}
    g["motor_pos"]["xyz"][:2] = np.clip(g["motor_pos"]["xyz"][:2], 0.005, None)
    quarter = g["body"]["w"] / 4.
    g["payload_pos"]["xy"] = np.clip(g["payload_pos"]["xy"], -quarter, quarter)
    g["arms_pos"]["angle"] = np.clip(g["arms_pos"]["angle"], 0., 90.)
    d["vel"] = np.clip(d["vel"], 0., 1.)
    d["omega_quadratic"] = np.clip(d["omega_quadratic"], 0., 1.)
    m["thrust_to_weight"] = np.clip(m["thrust_to_weight"], 1.2, None)
    m["torque_to_thrust"] = np.clip(m["torque_to_thrust"], 0.001, 1.)
    m["linearity"] = np.clip(m["linearity"], 0., 1.)
    m["assymetry"] = np.clip(m["assymetry"], 0.9, 1.1)
    for k in ("C_drag", "C_roll", "damp_time_up", "damp_time_down"):
        m[k] = np.clip(m[k], 0., None)
    if params_init is not None:                                           # keep propeller size consistent with t2w
        r0 = params_init["geom"]["propellers"]["r"]
        t2w, t2w0 = params_init["motor"]["thrust_to_weight"], m["thrust_to_weight"]
        g["propellers"]["r"] = r0 * (t2w / t2w0) ** 0.5
    return params


class Crazyflie(object):
    def sample(self, params=None):
        return quad_models.crazyflie_params()


class DefaultQuad(object):
    def sample(self, params=None):
        return quad_models.defaultquad_params()


class MediumQuad(object):
    def sample(self, params=None):
        return quad_models.mediumquad_params()


class CrazyflieLowInertia(object):
    def sample(self, params=None):
        return quad_models.crazyflie_lowinertia_params()


class ConstValueSampler(object):
    def __init__(self, params, params_change):
        self.params_change = copy.deepcopy(params_change)

    def sample(self, params):
        _update_existing(params, self.params_change)
        return params

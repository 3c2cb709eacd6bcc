"""Model samplers, dynamics randomisation and parameter sanity limits (host side).

Restates gym_art/quadrotor/quadrotor_randomization.py for the env-step path: check_quad_param_limits (:16-46),
get_dyn_randomization_params (:56-79), perturb_dyn_parameters (:82-111), resample_dyn_parameters (:113-147),
randomquad_parameters (:150-243), the sample_* helpers (:245-261) and the sampler classes (:329-377).  Every function
consumes np.random in the reference's order, so a seeded run reproduces the reference's parameter dicts draw for draw
(tests/test_host_params.py checks that against the live reference and against committed fixtures).  The sampled
dicts feed ModelConstants -> QsModel (one model per handle) or QuadVecEnv.set_models (one model per env, SURVEY.md
section 8 row f2).
"""
import copy

import numpy as np

from . import quad_models


def _walk(node, call):                                                    # quad_utils.py:16-21
    for key, item in node.items():
        if isinstance(item, dict):
            _walk(item, call)
        else:
            node[key] = call(key, item)


def _walk2(node1, node2, call):                                           # quad_utils.py:23-28
    for key, item in node1.items():
        if isinstance(item, dict):
            _walk2(item, node2[key], call)
        else:
            node1[key], node2[key] = call(key, item, node2[key])


def _update_existing(dst, upd):                                           # quad_utils.py:172-177
    for k, v in upd.items():
        if isinstance(dst[k], dict):
            _update_existing(dst[k], v)
        else:
            dst[k] = v


dict_update_existing = _update_existing


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


def get_dyn_randomization_params(quad_params, noise_ratio=0., noise_ratio_params=None):
    """:56-79 - a dict shaped like quad_params holding the relative noise ratio of every leaf."""
    noise_params = copy.deepcopy(quad_params)
    _walk(noise_params, lambda key, item: None if isinstance(item, str) else noise_ratio)
    if noise_ratio_params is not None:
        _update_existing(noise_params, noise_ratio_params)
    return noise_params


def perturb_dyn_parameters(params, noise_params, sampler="normal"):
    """:82-111 - sample around the nominal values; `ratio` is 2 sigma (normal) or the half width (uniform)."""
    def sample_normal(key, val, ratio):
        return np.random.normal(loc=val, scale=np.abs((ratio / 2) * np.array(val))), ratio

    def sample_uniform(key, val, ratio):
        val = np.array(val)
        return np.random.uniform(low=val - val * ratio, high=val + val * ratio), ratio

    params_new = copy.deepcopy(params)
    _walk2(params_new, noise_params, {"normal": sample_normal, "uniform": sample_uniform}[sampler])
    return check_quad_param_limits(params_new, params)


def resample_dyn_parameters(params, noise_params, sampler="uniform"):
    """:113-147 - leaves of noise_params are objects with .min / .max (absolute ranges)."""
    def sample_normal(key, val, mm):
        return np.random.normal(loc=(mm.min + mm.max) / 2, scale=(mm.max - mm.min) / 4)

    def sample_uniform(key, val, mm):
        return np.random.uniform(low=mm.min * np.ones_like(val), high=mm.max * np.ones_like(val))

    fn = {"normal": sample_normal, "uniform": sample_uniform}[sampler]
    params_new = copy.deepcopy(params)
    # the reference's walk_2dict assigns a 2-tuple from a 1-value callback here and would raise; the intent is clear
    _walk2(params_new, noise_params, lambda key, val, mm: (fn(key, val, mm), mm))
    return check_quad_param_limits(params_new, params)


def randomquad_parameters():
    """:150-243 - a random quadrotor between CrazyFlie and Hummingbird scale (densities, geometry, motors)."""
    g = {}
    dens = np.random.uniform(low=[500., 200., 500., 500., 200.], high=[2000., 2000., 2000., 4500., 300.])
    for k, part in enumerate(("body", "payload", "arms", "motors", "propellers")):
        g[part] = {"density": dens[k]}
    total_w = np.random.uniform(low=0.05, high=0.2)
    total_l = np.clip(np.random.normal(loc=1., scale=0.1), a_min=1.0, a_max=None) * total_w
    motor_z = np.random.normal(loc=0., scale=total_w / 8.)
    g["motor_pos"] = {"xyz": [total_w / 2., total_l / 2., motor_z]}
    g["motors"]["r"] = total_w * np.random.normal(loc=0.1, scale=0.01)
    g["motors"]["h"] = g["motors"]["r"] * np.random.normal(loc=1.0, scale=0.05)
    w_low, w_high = 0.25, 0.5
    w_coeff = np.random.uniform(low=w_low, high=w_high)
    g["body"]["w"] = w_coeff * total_w
    l_scale = (1. - (w_coeff - w_low) / (w_high - w_low))                 # narrower bodies are more elongated
    g["body"]["l"] = np.clip(np.random.normal(loc=1., scale=l_scale), a_min=1.0, a_max=None) * g["body"]["w"]
    g["body"]["h"] = np.random.uniform(low=0.1, high=1.5) * g["body"]["w"]
    pl = np.random.uniform(low=0.25, high=1.0, size=3)
    g["payload"]["w"] = pl[0] * g["body"]["w"]
    g["payload"]["l"] = pl[1] * g["body"]["l"]
    g["payload"]["h"] = pl[2] * g["body"]["h"]
    g["payload_pos"] = {"xy": np.random.normal(loc=0., scale=g["body"]["w"] / 10., size=2),
                        "z_sign": np.sign(np.random.uniform(low=-1, high=1))}
    g["arms"]["w"] = total_w * np.random.normal(loc=0.05, scale=0.005)
    g["arms"]["h"] = total_w * np.random.normal(loc=0.05, scale=0.005)
    g["arms_pos"] = {"angle": np.random.normal(loc=45., scale=10.), "z": motor_z - g["motors"]["h"] / 2.}
    t2w = np.random.uniform(low=1.5, high=3.5)
    g["propellers"]["h"] = 0.01
    g["propellers"]["r"] = 0.3 * total_w * (t2w / 2.0) ** 0.5
    damp = {"vel": 0.0, "omega_quadratic": 0.0}
    noise = {"thrust_noise_ratio": np.random.uniform(low=0.01, high=0.05)}
    damp_time_up = np.random.uniform(low=0.15, high=0.2)
    damp_time_down_scale = np.random.uniform(low=1.0, high=1.0)
    motor = {"thrust_to_weight": t2w, "torque_to_thrust": np.random.uniform(low=0.005, high=0.025),
             "assymetry": np.random.uniform(low=0.9, high=1.1, size=4), "linearity": 1.0, "C_drag": 0., "C_roll": 0.,
             "damp_time_up": damp_time_up, "damp_time_down": damp_time_down_scale * damp_time_up}
    return check_quad_param_limits({"geom": g, "damp": damp, "noise": noise, "motor": motor})


def sample_nodelay(params):                                               # :245-248
    params["motor"]["damp_time_up"] = 0.
    params["motor"]["damp_time_down"] = 0.
    return params


def sample_linearity(params):                                             # :250-252
    params["motor"]["linearity"] = np.random.uniform(low=0., high=1.)
    return params


def sample_t2w(params, t2w_min, t2w_max):                                 # :254-256
    params["motor"]["thrust_to_weight"] = np.random.uniform(low=t2w_min, high=t2w_max)
    return params


def sample_t2w_t2t(params, t2w_min, t2w_max, t2t_min=0.003, t2t_max=0.009):   # :258-261
    params["motor"]["thrust_to_weight"] = np.random.uniform(low=t2w_min, high=t2w_max)
    params["motor"]["torque_to_thrust"] = np.random.uniform(low=t2t_min, high=t2t_max)
    return params


def sample_simplified_random_dyn():
    """:263-326 - random parameters for the simplified (two rods) body: mass, arm length, motor parameters."""
    g = {"mass": np.random.uniform(low=0.020, high=0.6)}
    arm_length = np.random.uniform(low=0.05, high=0.5)
    g["arms"] = {"l": arm_length}
    mxy = arm_length * np.sqrt(2) / 4
    g["motor_pos"] = {"xyz": [mxy, mxy, 0.0]}
    t2w = np.random.uniform(low=1.5, high=3.5)
    damp = {"vel": 0.0, "omega_quadratic": 0.0}
    noise = {"thrust_noise_ratio": np.random.uniform(low=0.05, high=0.1)}
    damp_time_up = np.random.uniform(low=0.1, high=0.2)
    damp_time_down_scale = np.random.uniform(low=1.0, high=2.0)
    motor = {"thrust_to_weight": t2w, "torque_to_thrust": np.random.uniform(low=0.005, high=0.02),
             "assymetry": np.random.uniform(low=0.9, high=1.1, size=4), "linearity": 1.0, "C_drag": 0., "C_roll": 0.,
             "damp_time_up": damp_time_up, "damp_time_down": damp_time_down_scale * damp_time_up}
    return {"geom": g, "damp": damp, "noise": noise, "motor": motor}


class Crazyflie(object):                                                  # :329-343
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


class RandomQuad(object):
    def sample(self, params=None):
        return randomquad_parameters()


class RelativeSampler(object):                                            # :345-357
    def __init__(self, params, noise_ratio=0., noise_ratio_custom=None, sampler="normal"):
        self.noise_params = get_dyn_randomization_params(params, noise_ratio=noise_ratio,
                                                         noise_ratio_params=noise_ratio_custom)
        self.sampler = sampler

    def sample(self, params):
        return perturb_dyn_parameters(params=params, noise_params=self.noise_params, sampler=self.sampler)


class AbsoluteSampler(object):                                            # :359-369
    def __init__(self, params, noise_params, sampler="uniform"):
        self.noise_params = copy.deepcopy(noise_params)
        self.sampler = sampler

    def sample(self, params):
        return resample_dyn_parameters(params=params, noise_params=self.noise_params, sampler=self.sampler)


class ConstValueSampler(object):                                          # :371-377
    def __init__(self, params, params_change):
        self.params_change = copy.deepcopy(params_change)

    def sample(self, params):
        _update_existing(params, self.params_change)
        return params

"""Batched host pipeline for per-env dynamics randomisation: K random quadrotors -> the [QS_MODEL_ROWS, K, 4] model table
in one vectorised NumPy pass.

The scalar pipeline (quadrotor_randomization.randomquad_parameters -> inertia.QuadLink -> model.ModelConstants.rows)
costs ~0.4 ms per model, fine for hundreds of models but not for one model per env at 10^5..10^6 envs.  This module
restates the same three stages over arrays of K models: the sampling laws of quadrotor_randomization.py:150-243 (drawn
in vectorised order, so a seed gives the same distribution but not the same draws as K scalar calls), the link model
of inertia.py:182-310 and the derived constants / inverse Jacobian of quadrotor.py:142-208 and
quadrotor_control.py:192-211.  `params_of(batch, k)` converts one column back into the reference's nested dict;
tests/test_host_params.py checks that ModelConstants(params_of(batch, k)).rows(dt) equals column k of the table.
"""
import numpy as np

from .. import _lib
from .model import EPS, GRAV, OU_THETA, PROP_CCW
from .inertia import MOTOR_SIGNS


def sample_randomquad_batch(K, rng=None):
    """K draws of randomquad_parameters() (+ check_quad_param_limits) as a dict of arrays, leading dimension K."""
    r = np.random if rng is None else rng
    g = {}
    dens = r.uniform(low=[500., 200., 500., 500., 200.], high=[2000., 2000., 2000., 4500., 300.], size=(K, 5))
    total_w = r.uniform(0.05, 0.2, K)
    total_l = np.clip(r.normal(1., 0.1, K), 1.0, None) * total_w
    motor_z = r.normal(0., total_w / 8.)
    g["motor_xyz"] = np.stack([total_w / 2., total_l / 2., motor_z], axis=1)
    g["motors_r"] = total_w * r.normal(0.1, 0.01, K)
    g["motors_h"] = g["motors_r"] * r.normal(1.0, 0.05, K)
    w_low, w_high = 0.25, 0.5
    w_coeff = r.uniform(w_low, w_high, K)
    g["body_w"] = w_coeff * total_w
    l_scale = 1. - (w_coeff - w_low) / (w_high - w_low)
    g["body_l"] = np.clip(r.normal(1., l_scale), 1.0, None) * g["body_w"]
    g["body_h"] = r.uniform(0.1, 1.5, K) * g["body_w"]
    pl = r.uniform(0.25, 1.0, (K, 3))
    g["payload_w"], g["payload_l"], g["payload_h"] = pl[:, 0] * g["body_w"], pl[:, 1] * g["body_l"], pl[:, 2] * g["body_h"]
    g["payload_xy"] = r.normal(0., (g["body_w"] / 10.)[:, None], (K, 2))
    g["payload_z_sign"] = np.sign(r.uniform(-1, 1, K))
    g["arms_w"] = total_w * r.normal(0.05, 0.005, K)
    g["arms_h"] = total_w * r.normal(0.05, 0.005, K)
    g["arms_angle"] = r.normal(45., 10., K)
    g["arms_z"] = motor_z - g["motors_h"] / 2.
    t2w = r.uniform(1.5, 3.5, K)
    g["props_h"] = np.full(K, 0.01)
    g["props_r"] = 0.3 * total_w * (t2w / 2.0) ** 0.5
    for k, part in enumerate(("body", "payload", "arms", "motors", "props")):
        g[part + "_density"] = dens[:, k]
    damp_time_up = r.uniform(0.15, 0.2, K)
    b = dict(geom=g, thrust_to_weight=t2w, torque_to_thrust=r.uniform(0.005, 0.025, K), assymetry=r.uniform(0.9, 1.1, (K, 4)),
             linearity=np.ones(K), C_drag=np.zeros(K), C_roll=np.zeros(K), damp_time_up=damp_time_up,
             damp_time_down=r.uniform(1.0, 1.0, K) * damp_time_up, vel_damp=np.zeros(K), damp_omega_quadratic=np.zeros(K),
             thrust_noise_ratio=r.uniform(0.01, 0.05, K))
    return check_limits_batch(b)


def check_limits_batch(b):
    """check_quad_param_limits (quadrotor_randomization.py:16-46) over arrays."""
    g = b["geom"]
    for k in list(g):
        if k.split("_")[0] in ("body", "payload", "arms", "motors", "props") and k not in ("payload_xy", "payload_z_sign", "arms_angle", "arms_z"):
            g[k] = np.clip(g[k], 0., None)
    g["motor_xyz"][:, :2] = np.clip(g["motor_xyz"][:, :2], 0.005, None)
    q = (g["body_w"] / 4.)[:, None]
    g["payload_xy"] = np.clip(g["payload_xy"], -q, q)
    g["arms_angle"] = np.clip(g["arms_angle"], 0., 90.)
    b["vel_damp"] = np.clip(b["vel_damp"], 0., 1.)
    b["damp_omega_quadratic"] = np.clip(b["damp_omega_quadratic"], 0., 1.)
    b["thrust_to_weight"] = np.clip(b["thrust_to_weight"], 1.2, None)
    b["torque_to_thrust"] = np.clip(b["torque_to_thrust"], 0.001, 1.)
    b["linearity"] = np.clip(b["linearity"], 0., 1.)
    b["assymetry"] = np.clip(b["assymetry"], 0.9, 1.1)
    for k in ("C_drag", "C_roll", "damp_time_up", "damp_time_down"):
        b[k] = np.clip(b[k], 0., None)
    return b


def params_of(b, k):
    """Column k as the reference's nested parameter dict (what the scalar pipeline consumes)."""
    g = b["geom"]
    geom = {
        "body": {"density": g["body_density"][k], "w": g["body_w"][k], "l": g["body_l"][k], "h": g["body_h"][k]},
        "payload": {"density": g["payload_density"][k], "w": g["payload_w"][k], "l": g["payload_l"][k], "h": g["payload_h"][k]},
        "arms": {"density": g["arms_density"][k], "w": g["arms_w"][k], "h": g["arms_h"][k]},
        "motors": {"density": g["motors_density"][k], "r": g["motors_r"][k], "h": g["motors_h"][k]},
        "propellers": {"density": g["props_density"][k], "h": g["props_h"][k], "r": g["props_r"][k]},
        "motor_pos": {"xyz": list(g["motor_xyz"][k])},
        "payload_pos": {"xy": np.array(g["payload_xy"][k]), "z_sign": g["payload_z_sign"][k]},
        "arms_pos": {"angle": g["arms_angle"][k], "z": g["arms_z"][k]},
    }
    motor = {"thrust_to_weight": b["thrust_to_weight"][k], "torque_to_thrust": b["torque_to_thrust"][k],
             "assymetry": np.array(b["assymetry"][k]), "linearity": b["linearity"][k], "C_drag": b["C_drag"][k],
             "C_roll": b["C_roll"][k], "damp_time_up": b["damp_time_up"][k], "damp_time_down": b["damp_time_down"][k]}
    return {"geom": geom, "damp": {"vel": b["vel_damp"][k], "omega_quadratic": b["damp_omega_quadratic"][k]},
            "noise": {"thrust_noise_ratio": b["thrust_noise_ratio"][k]}, "motor": motor}


def quadlink_batch(g):
    """inertia.py:182-310 over K models: total mass [K], diagonal of the inertia about the COM [K,3], motor positions
    relative to the COM [K,4,3], motor_xyz [K,3]."""
    K = g["body_w"].shape[0]
    ang = g["arms_angle"] / 180. * np.pi
    ang = np.where(ang == 0., 0.01, ang)
    mxyz = g["motor_xyz"]
    dy = mxyz[:, 1] - g["body_w"] / 2.
    arm_l = dy / np.sin(ang)
    arm_xyz = np.stack([mxyz[:, 0] - dy / (2 * np.tan(ang)), mxyz[:, 1] - dy / 2, g["arms_z"]], axis=1)
    sign = MOTOR_SIGNS.T[None]                                             # [1,4,3]
    motors_xyz = sign * mxyz[:, None, :]
    props_xyz = motors_xyz.copy()
    props_xyz[:, :, 2] += (g["motors_h"] / 2. + g["props_h"])[:, None]
    arms_xyz = sign * arm_xyz[:, None, :]
    payload_xyz = np.concatenate([g["payload_xy"], (g["payload_z_sign"] * (g["body_h"] + g["payload_h"]) / 2)[:, None]], axis=1)

    def box(l, w, h, dens):
        m = dens * l * w * h
        return m, np.stack([m * (h ** 2 + w ** 2), m * (l ** 2 + h ** 2), m * (w ** 2 + l ** 2)], axis=1) / 12.0

    def cyl(h, r, dens):
        m = dens * np.pi * h * r ** 2
        side = m * (3 * r ** 2 + h ** 2) / 12.0
        return m, np.stack([side, side, 0.5 * m * r ** 2], axis=1)

    mb, Ib = box(g["body_l"], g["body_w"], g["body_h"], g["body_density"])
    mp, Ip = box(g["payload_l"], g["payload_w"], g["payload_h"], g["payload_density"])
    ma, Ia = box(arm_l, g["arms_w"], g["arms_h"], g["arms_density"])
    mm, Im = cyl(g["motors_h"], g["motors_r"], g["motors_density"])
    mr, Ir = cyl(g["props_h"], g["props_r"], g["props_density"])
    # arms are yawed by -+angle: the diagonal of Rz D Rz^T
    c2, s2 = np.cos(ang) ** 2, np.sin(ang) ** 2
    Ia_rot = np.stack([c2 * Ia[:, 0] + s2 * Ia[:, 1], s2 * Ia[:, 0] + c2 * Ia[:, 1], Ia[:, 2]], axis=1)
    mass = mb + mp + 4 * (ma + mm + mr)
    com = (mp[:, None] * payload_xyz + ma[:, None] * arms_xyz.sum(1) + mm[:, None] * motors_xyz.sum(1) +
           mr[:, None] * props_xyz.sum(1)) / mass[:, None]                 # the body sits at the origin

    def shifted(I, m, xyz):                                               # parallel axes, diagonal only
        d = xyz - (com if xyz.ndim == 2 else com[:, None, :])
        x, y, z = d[..., 0], d[..., 1], d[..., 2]
        add = np.stack([y * y + z * z, x * x + z * z, x * x + y * y], axis=-1)
        if xyz.ndim == 3:
            return (I[:, None, :] + m[:, None, None] * add).sum(1)
        return I + m[:, None] * add

    inertia = (shifted(Ib, mb, np.zeros((K, 3))) + shifted(Ip, mp, payload_xyz) + shifted(Ia_rot, ma, arms_xyz) +
               shifted(Im, mm, motors_xyz) + shifted(Ir, mr, props_xyz))
    return mass, inertia, motors_xyz - com[:, None, :], mxyz


def model_rows_batch(b, dt):
    """[QS_MODEL_ROWS, K, 4] model table (row layout in include/quadsim.h) for the K models of `b`."""
    mass, inertia, prop_pos, mxyz = quadlink_batch(b["geom"])
    K = mass.shape[0]
    asym = b["assymetry"] * 4. / np.sum(b["assymetry"], axis=1, keepdims=True)               # quadrotor.py:174
    thrust_max = GRAV * (mass * b["thrust_to_weight"])[:, None] * asym / 4.0                  # :175
    torque_max = b["torque_to_thrust"][:, None] * thrust_max                                  # :176
    cross = np.cross(prop_pos, np.array([0., 0., 1.]))                                        # :182  [K,4,3]
    torque = thrust_max[:, None, :] * np.transpose(cross, (0, 2, 1))                          # [K,3,4]
    torque[:, 2, :] = torque_max * PROP_CCW
    J = np.concatenate([(thrust_max / mass[:, None])[:, None, :], torque / inertia[:, :, None]], axis=1)   # [K,4,4]
    Jinv = np.linalg.inv(J)
    r = np.zeros((_lib.QS_MODEL_ROWS, K, 4))
    r[0] = thrust_max
    r[1] = torque_max * PROP_CCW
    r[2] = prop_pos[:, :, 1]
    r[3] = -prop_pos[:, :, 0]
    r[4] = prop_pos[:, :, 2]
    r[5, :, :3], r[5, :, 3] = inertia, mass
    r[6, :, :3], r[6, :, 3] = 1.0 / inertia, 1.0 / mass
    r[7, :, 0] = np.minimum(1.0, 4.0 * dt / (b["damp_time_up"] + EPS))
    r[7, :, 1] = np.minimum(1.0, 4.0 * dt / (b["damp_time_down"] + EPS))
    r[7, :, 2] = 0.2 * b["thrust_noise_ratio"]
    r[7, :, 3] = np.linalg.norm(mxyz[:, :2], axis=1)
    r[8] = np.stack([b["linearity"], b["C_drag"], b["C_roll"], b["thrust_to_weight"]], axis=1)
    r[9, :, 0], r[9, :, 1], r[9, :, 2] = b["vel_damp"], b["damp_omega_quadratic"], b["torque_to_thrust"]
    r[10:14] = np.transpose(Jinv, (1, 0, 2))
    return r


__all__ = ["sample_randomquad_batch", "check_limits_batch", "params_of", "quadlink_batch", "model_rows_batch", "OU_THETA"]

"""TEST INFRASTRUCTURE ONLY - CPU oracle for the quadrotor env-step hot path.

A batched float64 NumPy restatement of the reference algorithm (amolchanov86/gym_art).  It is the
checker for the CUDA path (tests/, __graft_entry__.smoke(), bench.py's cpu_baseline leg) and is
never imported by anything under gym_art_b200/.  Every function cites the reference file:line it
follows (paths relative to the reference root, gym_art/quadrotor/...).

Pinning: the reference's own tests hold no golden vectors for this path, so the
oracle is pinned against (i) the live reference imported under stubs in the build container
(tests/test_oracle_vs_reference.py) and (ii) fixtures generated from the live reference by
oracle/gen_golden.py and committed under tests/golden/ (tests/test_oracle_golden.py).
The swarm collision stage (collide_swarm) has NO reference counterpart: parity unpinned, spec is ours.

All randomness is *injected*: callers pass unit normals / unit uniforms in the slot layout of
include/quadsim.h (QS_OBS_Z_*, QS_RESET_Z_*).  Axis 0 is always the env index.
"""
import numpy as np

GRAV = 9.81   # quadrotor.py:53
EPS = 1e-6    # quadrotor.py:54

# injected-noise slot layout, mirrored by include/quadsim.h ----------------------------------------
OBS_Z = 33            # n_pos3 u_pos3 n_vel3 u_vel3 n_gyro3 n_gyro_rw3 n_theta3 u_theta3 n_accS3 n_accD3 | n_t2w n_t2t (unused)
RESET_Z = 18 + OBS_Z  # u_pos3 yaw1 u_vel3 u_velmag1 u_om3 u_ommag1 n_up3 n_fwd3 | obs_z for the post-reset obs
GOAL_Z = 4            # excite goal u_xy2 u_z1 (quadrotor.py:957-962) | resample_goal u_z1 (:1078-1079)
(Z_NPOS, Z_UPOS, Z_NVEL, Z_UVEL, Z_NGYRO, Z_NGYRO_RW, Z_NTHETA, Z_UTHETA, Z_NACC_S, Z_NACC_D) = range(0, 30, 3)
Z_NT2W, Z_NT2T = 30, 31   # unit normals of the thrust-to-weight / torque-to-thrust observation noise (get_state.py:334, :379)
INIT_DEFAULT, INIT_RANDOM, INIT_PITCH_ROLL_RESTRICTED = 0, 1, 2   # which law the reset draws (vel, rot, omega) from
(RZ_UPOS, RZ_YAW, RZ_UVEL, RZ_UVELMAG, RZ_UOM, RZ_UOMMAG, RZ_NUP, RZ_NFWD, RZ_OBS) = (0, 3, 4, 7, 8, 11, 12, 15, 18)

CTRL_RAW, CTRL_VERTICAL, CTRL_VERTPLANE, CTRL_SHIFTED, CTRL_OMEGA_THRUST, CTRL_VELOCITY_YAW, CTRL_MELLINGER = range(7)
CTRL_ACT_DIM = {CTRL_RAW: 4, CTRL_VERTICAL: 1, CTRL_VERTPLANE: 2, CTRL_SHIFTED: 4, CTRL_OMEGA_THRUST: 4,
                CTRL_VELOCITY_YAW: 4, CTRL_MELLINGER: 4}

OBS_REPRS = {  # get_state.py:5,134,147,219,236,249 -> (relative frame, +h, +acc, +act); the rest of the table below
    "xyz_vxyz_R_omega": (False, False, False, False),
    "xyz_vxyz_R_omega_h": (False, True, False, False),
    "xyzr_vxyzr_R_omega": (True, False, False, False),
    "xyzr_vxyzr_R_omega_h": (True, True, False, False),
    "xyz_vxyz_R_omega_acc_act": (False, False, True, True),
    "xyz_vxyz_R_omega_act": (False, False, False, True),
    # builders that only lack module-level names in the reference (`normal`, `R2quat`: get_state.py has no imports for
    # them); oracle/ref_shim.py supplies the names, the reference code itself is untouched
    "xyz_vxyz_quat_omega": (False, False, False, False),                 # :276
    "xyzr_vxyzr_quat_omega": (True, False, False, False),                # :289
    "xyzr_vxyzr_quat_omega_h": (True, True, False, False),               # :307
    "xyzr_vxyzr_R_omega_t2w": (True, False, False, False),               # :325
    "xyz_vxyz_R_omega_t2w": (False, False, False, False),                # :345
    "xyz_vxyz_R_omega_t2w_t2t": (False, False, False, False),            # :363
}
T2W_MIN, T2W_MAX, T2T_MIN, T2T_MAX = 1.5, 10.0, 0.005, 1.0              # quadrotor.py:707-712


def obs_layout(obs_repr):
    """(relative frame, +h, +acc, +act, quaternion instead of R, +t2w, +t2t)"""
    rel, h, acc, act = OBS_REPRS[obs_repr]
    comps = obs_repr.split("_")
    return rel, h, acc, act, "quat" in comps, "t2w" in comps, "t2t" in comps


def obs_dim(obs_repr):
    rel, h, acc, act, quat, t2w, t2t = obs_layout(obs_repr)
    return 9 + (4 if quat else 9) + int(h) + 3 * int(acc) + 4 * int(act) + int(t2w) + int(t2t)


# ==================================================================================================
# a0  host-side model constants
# ==================================================================================================
def _box_I(l, w, h, m):          # inertia.py:88-95 (BoxLink.I_com)
    return np.diag([m * (h ** 2 + w ** 2) / 12.0, m * (l ** 2 + h ** 2) / 12.0, m * (w ** 2 + l ** 2) / 12.0])


def _cyl_I(h, r, m):             # inertia.py:151-157 (CylinderLink.I_com)
    return np.diag([m * (3 * r ** 2 + h ** 2) / 12.0, m * (3 * r ** 2 + h ** 2) / 12.0, 0.5 * m * r ** 2])


def _translate_I(I, m, xyz):     # inertia.py:23-37 (incl. the xz typo at :34; only the diagonal is consumed)
    x, y, z = xyz
    out = np.zeros((3, 3))
    out[0, 0] = I[0, 0] + m * (y ** 2 + z ** 2)
    out[1, 1] = I[1, 1] + m * (x ** 2 + z ** 2)
    out[2, 2] = I[2, 2] + m * (x ** 2 + y ** 2)
    out[0, 1] = out[1, 0] = I[0, 1] + m * x * y
    out[0, 2] = out[2, 0] = I[0, 1] + m * x * z
    out[1, 2] = out[2, 1] = I[1, 2] + m * y * z
    return out


def quadlink_constants(geom):
    """inertia.py:182-310 (QuadLink): total mass, inertia about the COM, propeller positions."""
    body, payload, arms, motors, props = (geom[k] for k in ("body", "payload", "arms", "motors", "propellers"))
    arm_angle = float(geom["arms_pos"]["angle"]) / 180.0 * np.pi          # :215
    if arm_angle == 0.0:
        arm_angle = 0.01
    motor_xyz = np.array(geom["motor_pos"]["xyz"], dtype=np.float64)      # :218
    delta_y = motor_xyz[1] - body["w"] / 2.0                              # :219
    arm_l = arms["l"] if "l" in arms else delta_y / np.sin(arm_angle)      # :220-224
    arm_xyz = np.array([motor_xyz[0] - delta_y / (2 * np.tan(arm_angle)), motor_xyz[1] - delta_y / 2,
                        geom["arms_pos"]["z"]])                           # :228-230
    sx = np.array([1.0, -1.0, -1.0, 1.0])                                 # :236 FR, BR, BL, FL
    sy = np.array([-1.0, -1.0, 1.0, 1.0])                                 # :237
    sign = np.stack([sx, sy, np.ones(4)])                                 # :238
    motors_coord = sign * motor_xyz[:, None]                              # :239
    props_coord = motors_coord.copy()
    props_coord[2] = props_coord[2] + motors["h"] / 2.0 + props["h"]      # :241
    arm_angles = [-arm_angle, arm_angle, -arm_angle, arm_angle]           # :242-246
    arms_coord = sign * arm_xyz[:, None]                                  # :247

    masses, inertias, rots, xyzs = [], [], [], []

    def add(m, I, xyz, alpha=None):
        masses.append(float(m))
        inertias.append(I)
        xyzs.append(np.array(xyz, dtype=np.float64))
        if alpha:                                                         # LinkPose :170-177
            rots.append(np.array([[np.cos(alpha), -np.sin(alpha), 0.0], [np.sin(alpha), np.cos(alpha), 0.0],
                                  [0.0, 0.0, 1.0]]))
        else:
            rots.append(np.eye(3))

    def box_m(d, l):             # BoxLink.__init__/compute_m, inertia.py:75-82, :98-99
        return d["m"] if d.get("m") is not None else d["density"] * l * d["w"] * d["h"]

    def cyl_m(d):                # CylinderLink.__init__/compute_m, inertia.py:139-146, :160-161
        return d["m"] if d.get("m") is not None else d["density"] * np.pi * d["h"] * d["r"] ** 2

    mb, mp, ma, mm, mr = box_m(body, body["l"]), box_m(payload, payload["l"]), box_m(arms, arm_l), cyl_m(motors), cyl_m(props)
    add(mb, _box_I(body["l"], body["w"], body["h"], mb), [0.0, 0.0, 0.0])
    pz = np.sign(geom["payload_pos"]["z_sign"]) * (body["h"] + payload["h"]) / 2       # :268
    add(mp, _box_I(payload["l"], payload["w"], payload["h"], mp), list(geom["payload_pos"]["xy"]) + [pz])
    for i in range(4):
        add(ma, _box_I(arm_l, arms["w"], arms["h"], ma), arms_coord[:, i], alpha=arm_angles[i])
    for i in range(4):
        add(mm, _cyl_I(motors["h"], motors["r"], mm), motors_coord[:, i])
    for i in range(4):
        add(mr, _cyl_I(props["h"], props["r"], mr), props_coord[:, i])

    m_total = np.sum(masses)                                              # :308-310
    com = sum(masses[i] * xyzs[i] for i in range(len(masses))) / m_total  # :281-282
    xyzs = [x - com for x in xyzs]                                        # :287-288
    I_com = sum(_translate_I(rots[i] @ inertias[i] @ rots[i].T, masses[i], xyzs[i])
                for i in range(len(masses)))                              # :300-306
    prop_pos = np.array(xyzs[6:10])                                       # :309 (motor poses, despite the name)
    return dict(mass=m_total, I_com=I_com, prop_pos=prop_pos, motor_xyz=motor_xyz)


def quadlink_simplified_constants(geom):
    """inertia.py:312-420 (QuadLinkSimplified): two rods of half the mass each, massless propellers."""
    defaults = {"body": {"l": 0.03, "w": 0.03, "h": 0.004, "m": 0.005}, "payload": {"l": 0.035, "w": 0.02, "h": 0.008, "m": 0.01},
                "arms": {"w": 0.005, "h": 0.005, "m": 0.001}, "motors": {"h": 0.02, "r": 0.0035, "m": 0.0015},
                "arms_pos": {"angle": 45.0, "z": 0.0}}                    # :327-336 (CrazyFlie values for unspecified parts)
    geom = dict(defaults, **geom)
    arm_length = np.sqrt(geom["motor_pos"]["xyz"][0] ** 2 * 2) * 2        # :345
    mxy = arm_length * np.sqrt(2) / 4                                     # :347
    if "mass" in geom:
        mass = geom["mass"]
    else:                                                                 # :351-358, propellers already massless (:346)
        def bm(d):
            return d["m"] if d.get("m") is not None else d["density"] * d["l"] * d["w"] * d["h"]

        def cm(d):
            return d["m"] if d.get("m") is not None else d["density"] * np.pi * d["h"] * d["r"] ** 2
        mass = bm(geom["body"]) + bm(geom["payload"]) + 4 * bm(geom["arms"]) + 4 * cm(geom["motors"])
    rod_m = mass / 2                                                      # :360
    ang = float(geom["arms_pos"]["angle"]) / 180.0 * np.pi
    if ang == 0.0:
        ang = 0.01
    a = rod_m * arm_length ** 2 / 12.0                                    # RodLink.I_com :121-126
    I = np.zeros((3, 3))
    for al in (-ang, ang):                                                # :391-394, rods at the COM
        Rz = np.array([[np.cos(al), -np.sin(al), 0.0], [np.sin(al), np.cos(al), 0.0], [0.0, 0.0, 1.0]])
        I = I + Rz @ np.diag([a, 0.0, a]) @ Rz.T
    sx = np.array([1.0, -1.0, -1.0, 1.0])
    sy = np.array([-1.0, -1.0, 1.0, 1.0])
    z_arm = geom["arms_pos"]["z"]
    prop_pos = np.stack([sx * mxy, sy * mxy, np.zeros(4) - z_arm], axis=1)   # motors at z = 0, COM at the rods' height
    return dict(mass=2 * rod_m, I_com=I, prop_pos=prop_pos, motor_xyz=np.array([mxy, mxy, 0.0]))


def model_constants(params, simplified=False):
    """quadrotor.py:77-208 (QuadrotorDynamics.__init__/update_model) -> flat dict of float64 constants."""
    ql = quadlink_simplified_constants(params["geom"]) if simplified else quadlink_constants(params["geom"])
    motor, damp, noise = params["motor"], params["damp"], params["noise"]
    mass = ql["mass"]
    inertia = np.diagonal(ql["I_com"]).copy()                             # :152
    asym = np.array(motor.get("assymetry", [1.0, 1.0, 1.0, 1.0]), dtype=np.float64)
    asym = asym * 4.0 / np.sum(asym)                                      # :174
    t2w = float(motor["thrust_to_weight"])
    t2t = float(motor["torque_to_thrust"])
    thrust_max = GRAV * mass * t2w * asym / 4.0                           # :175
    torque_max = t2t * thrust_max                                         # :176
    prop_pos = ql["prop_pos"]
    prop_cross = np.cross(prop_pos, [0.0, 0.0, 1.0])                      # :182
    m = dict(
        mass=mass, inertia=inertia, thrust_to_weight=t2w, torque_to_thrust=t2t,
        thrust_max=thrust_max, torque_max=torque_max, prop_pos=prop_pos, prop_cross=prop_cross,
        prop_ccw=np.array([-1.0, 1.0, -1.0, 1.0]),                        # :88
        linearity=float(motor["linearity"]), C_drag=float(motor["C_drag"]), C_roll=float(motor["C_roll"]),
        damp_time_up=float(motor["damp_time_up"]), damp_time_down=float(motor["damp_time_down"]),
        vel_damp=float(damp["vel"]), damp_omega_quadratic=float(damp["omega_quadratic"]),
        thrust_noise_ratio=float(noise["thrust_noise_ratio"]),
        ou_sigma=0.2 * float(noise["thrust_noise_ratio"]), ou_theta=0.15,  # :198, quad_utils.py:181
        arm=float(np.linalg.norm(ql["motor_xyz"][:2])),                   # :200
        omega_max=40.0,                                                   # :91
        gravity=GRAV,                                                     # :93 (self.gravity; accelerometer only)
    )
    m["Jinv"] = jacobian_inverse(m)
    return m


def jacobian(m):
    """quadrotor_control.py:192-203."""
    torque = m["thrust_max"] * m["prop_cross"].T                          # [3,4]
    torque[2, :] = m["torque_max"] * m["prop_ccw"]
    thrust = m["thrust_max"] * np.ones((1, 4))
    dw = (1.0 / m["inertia"])[:, None] * torque
    dv = thrust / m["mass"]
    return np.vstack([dv, dw])


def jacobian_inverse(m):
    return np.linalg.inv(jacobian(m))                                     # quadrotor_control.py:211


# ==================================================================================================
# small batched helpers (quad_utils.py)
# ==================================================================================================
def _cross(a, b):                # quad_utils.py:31
    return np.stack([a[:, 1] * b[:, 2] - a[:, 2] * b[:, 1], a[:, 2] * b[:, 0] - a[:, 0] * b[:, 2],
                     a[:, 0] * b[:, 1] - a[:, 1] * b[:, 0]], axis=1)


def _norm3(x):
    return (x[:, 0] ** 2 + x[:, 1] ** 2 + x[:, 2] ** 2) ** 0.5


def _normalize(x):               # quad_utils.py:35-41: vectors shorter than 1e-5 are returned unscaled
    n = _norm3(x)
    small = n < 0.00001
    return np.where(small[:, None], x, x / np.where(small, 1.0, n)[:, None])


def _clamp_norm(x, maxnorm):     # quad_utils.py:63-67
    n = _norm3(x)
    over = n > maxnorm
    return np.where(over[:, None], (maxnorm / np.where(over, n, 1.0))[:, None] * x, x)


def quat2R(q):                   # quad_utils.py:82-87
    qw, qx, qy, qz = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    R = np.empty((q.shape[0], 3, 3))
    R[:, 0, 0] = 1.0 - 2 * qy ** 2 - 2 * qz ** 2
    R[:, 0, 1] = 2 * qx * qy - 2 * qz * qw
    R[:, 0, 2] = 2 * qx * qz + 2 * qy * qw
    R[:, 1, 0] = 2 * qx * qy + 2 * qz * qw
    R[:, 1, 1] = 1.0 - 2 * qx ** 2 - 2 * qz ** 2
    R[:, 1, 2] = 2 * qy * qz - 2 * qx * qw
    R[:, 2, 0] = 2 * qx * qz - 2 * qy * qw
    R[:, 2, 1] = 2 * qy * qz + 2 * qx * qw
    R[:, 2, 2] = 1.0 - 2 * qx ** 2 - 2 * qy ** 2
    return R


def quatXquat(q, t):             # quad_utils.py:92-99
    out = np.empty_like(q)
    out[:, 0] = q[:, 0] * t[:, 0] - q[:, 1] * t[:, 1] - q[:, 2] * t[:, 2] - q[:, 3] * t[:, 3]
    out[:, 1] = q[:, 0] * t[:, 1] + q[:, 1] * t[:, 0] - q[:, 2] * t[:, 3] + q[:, 3] * t[:, 2]
    out[:, 2] = q[:, 0] * t[:, 2] + q[:, 1] * t[:, 3] + q[:, 2] * t[:, 0] - q[:, 3] * t[:, 1]
    out[:, 3] = q[:, 0] * t[:, 3] - q[:, 1] * t[:, 2] + q[:, 2] * t[:, 1] + q[:, 3] * t[:, 0]
    return out


def rot2quat(R):                 # sensor_noise.py:26-55 (four-branch)
    n = R.shape[0]
    q = np.empty((n, 4))
    tr = R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]
    b0 = tr > 0
    b1 = ~b0 & (R[:, 0, 0] > R[:, 1, 1]) & (R[:, 0, 0] > R[:, 2, 2])
    b2 = ~b0 & ~b1 & (R[:, 1, 1] > R[:, 2, 2])
    b3 = ~b0 & ~b1 & ~b2
    with np.errstate(invalid="ignore", divide="ignore"):
        S = (tr + 1.0) ** 0.5 * 2
        q0 = np.stack([0.25 * S, (R[:, 2, 1] - R[:, 1, 2]) / S, (R[:, 0, 2] - R[:, 2, 0]) / S,
                       (R[:, 1, 0] - R[:, 0, 1]) / S], axis=1)
        S = (1.0 + R[:, 0, 0] - R[:, 1, 1] - R[:, 2, 2]) ** 0.5 * 2
        q1 = np.stack([(R[:, 2, 1] - R[:, 1, 2]) / S, 0.25 * S, (R[:, 0, 1] + R[:, 1, 0]) / S,
                       (R[:, 0, 2] + R[:, 2, 0]) / S], axis=1)
        S = (1.0 + R[:, 1, 1] - R[:, 0, 0] - R[:, 2, 2]) ** 0.5 * 2
        q2 = np.stack([(R[:, 0, 2] - R[:, 2, 0]) / S, (R[:, 0, 1] + R[:, 1, 0]) / S, 0.25 * S,
                       (R[:, 1, 2] + R[:, 2, 1]) / S], axis=1)
        S = (1.0 + R[:, 2, 2] - R[:, 0, 0] - R[:, 1, 1]) ** 0.5 * 2
        q3 = np.stack([(R[:, 1, 0] - R[:, 0, 1]) / S, (R[:, 0, 2] + R[:, 2, 0]) / S,
                       (R[:, 1, 2] + R[:, 2, 1]) / S, 0.25 * S], axis=1)
    q[b0] = q0[b0]
    q[b1] = q1[b1]
    q[b2] = q2[b2]
    q[b3] = q3[b3]
    return q


def R2quat(R):                   # quad_utils.py:101-109 (single branch: needs 1 + trace > 0, as the reference)
    w = np.sqrt(1.0 + R[:, 0, 0] + R[:, 1, 1] + R[:, 2, 2]) / 2.0
    w4 = 4.0 * w
    return np.stack([w, (R[:, 2, 1] - R[:, 1, 2]) / w4, (R[:, 0, 2] - R[:, 2, 0]) / w4, (R[:, 1, 0] - R[:, 0, 1]) / w4], axis=1)


def euler2mat_sxyz(ai, aj, ak):
    """transforms3d.euler.euler2mat(ai, aj, ak) with its default axes 'sxyz' (transforms3d is not installed here; this
    is its documented convention: static-frame rotations about x, then y, then z, R = Rz(ak) Ry(aj) Rx(ai)).  Used by
    QuadrotorDynamics.pitch_roll_restricted_random_state, quadrotor.py:257."""
    ci, si, cj, sj, ck, sk = np.cos(ai), np.sin(ai), np.cos(aj), np.sin(aj), np.cos(ak), np.sin(ak)
    R = np.empty(np.shape(ai) + (3, 3))
    R[..., 0, 0], R[..., 0, 1], R[..., 0, 2] = cj * ck, sj * si * ck - ci * sk, sj * ci * ck + si * sk
    R[..., 1, 0], R[..., 1, 1], R[..., 1, 2] = cj * sk, sj * si * sk + ci * ck, sj * ci * sk - si * ck
    R[..., 2, 0], R[..., 2, 1], R[..., 2, 2] = -sj, cj * si, cj * ci
    return R


def quat_from_small_angle(theta):  # sensor_noise.py:9-21
    q_squared = (theta[:, 0] ** 2 + theta[:, 1] ** 2 + theta[:, 2] ** 2) / 4.0
    lt = q_squared < 1
    w_lt = np.sqrt(np.where(lt, 1 - q_squared, 0.0))
    w_ge = 1.0 / (1 + q_squared) ** 0.5
    w = np.where(lt, w_lt, w_ge)
    f = np.where(lt, 0.5, 0.5 * w_ge)
    q = np.concatenate([w[:, None], theta * f[:, None]], axis=1)
    return q / np.sqrt(np.sum(q ** 2, axis=1))[:, None]


def polar_orthonormalize(R):     # quadrotor.py:384-385  (u, s, v = svd(R); R = u @ v)
    u, s, vt = np.linalg.svd(R)
    return u @ vt


# ==================================================================================================
# a6  SensorNoise (sensor_noise.py:57-170)
# ==================================================================================================
DEFAULT_SENSE_NOISE = dict(pos_norm_std=0.005, pos_unif_range=0.0, vel_norm_std=0.01, vel_unif_range=0.0,
                           quat_norm_std=0.0, quat_unif_range=0.0, gyro_norm_std=0.0, gyro_noise_density=0.000175,
                           gyro_random_walk=0.0105, gyro_bias_correlation_time=1000.0,
                           acc_static_noise_std=0.002, acc_dynamic_noise_ratio=0.005)   # sensor_noise.py:58-63


def gyro_bias_coeffs(sn, dt):    # sensor_noise.py:165-167
    sigma_g_d = sn["gyro_noise_density"] / (dt ** 0.5)
    T = sn["gyro_bias_correlation_time"]
    sigma_b_g_d = (-(sigma_g_d ** 2) * (T / 2) * (np.exp(-2 * dt / T) - 1)) ** 0.5
    pi_g_d = np.exp(-dt / T)
    return pi_g_d, sigma_b_g_d


def add_noise(sn, pos, vel, rot, omega, acc, dt, z, gyro_bias, n_calls=1):
    """sensor_noise.py:100-159.  sn=None is bypass.  z: [N, OBS_Z] unit draws.  rot: [N,3,3] rotation matrices, or
    [N,4] quaternions (the quat observation builders pass R2quat(rot), sensor_noise.py:148-151).

    n_calls: how many times the reference invokes add_noise around this observable one
    (3 per env step - quadrotor.py:947,970,988 - and 1 at reset, :1143).  It only matters for the
    RotorS gyro-bias walk, whose state advances on every call (sensor_noise.py:169): the n_calls-1
    unobserved updates are folded in exactly, b <- pi^n b + sigma_b * s_n * z with
    s_n = sqrt(sum_k pi^(2k)), z the single effective unit normal supplied in Z_NGYRO.
    Returns (pos, vel, rot, omega, acc, new_gyro_bias).
    """
    if sn is None:
        return pos, vel, rot, omega, acc, gyro_bias
    npos = pos + sn["pos_norm_std"] * z[:, Z_NPOS:Z_NPOS + 3] + sn["pos_unif_range"] * z[:, Z_UPOS:Z_UPOS + 3]
    nvel = vel + sn["vel_norm_std"] * z[:, Z_NVEL:Z_NVEL + 3] + sn["vel_unif_range"] * z[:, Z_UVEL:Z_UVEL + 3]
    if sn["gyro_norm_std"] != 0.0:                                        # :127-128, :162-170
        pi1, sb1 = gyro_bias_coeffs(sn, dt)
        pin = pi1 ** n_calls
        sn_eff = sb1 * np.sqrt(sum(pi1 ** (2 * k) for k in range(n_calls)))
        gyro_bias = pin * gyro_bias + sn_eff * z[:, Z_NGYRO:Z_NGYRO + 3]
        nomega = omega + gyro_bias + sn["gyro_random_walk"] * z[:, Z_NGYRO_RW:Z_NGYRO_RW + 3]
    else:
        nomega = omega + sn["gyro_noise_density"] * z[:, Z_NGYRO:Z_NGYRO + 3]          # :130-131
    theta = sn["quat_norm_std"] * z[:, Z_NTHETA:Z_NTHETA + 3] + sn["quat_unif_range"] * z[:, Z_UTHETA:Z_UTHETA + 3]
    if rot.ndim == 2:
        nrot = quatXquat(rot, quat_from_small_angle(theta))                           # :148-151
    else:
        nrot = quat2R(quatXquat(rot2quat(rot), quat_from_small_angle(theta)))         # :144-147
    nacc = acc + sn["acc_static_noise_std"] * z[:, Z_NACC_S:Z_NACC_S + 3] \
        + acc * (sn["acc_dynamic_noise_ratio"] * z[:, Z_NACC_D:Z_NACC_D + 3])          # :156-157
    return npos, nvel, nrot, nomega, nacc, gyro_bias


# ==================================================================================================
# affine dynamics representation s' = dt (F(s) + G(s) u)   (quadrotor.py:457-518)
# ==================================================================================================
CONTROL_MX = {"1D": np.ones((4, 1)), "2D": np.array([[1.0, 0.0], [1.0, 0.0], [0.0, 1.0], [0.0, 1.0]]), "3D": np.eye(4)}   # :121-129


def affine_F(m, s, dt):
    """quadrotor.py:461-497.  s: [N, 21] = xyz, Vxyz, R (row-major), omega, goal.  Returns [N, 21]."""
    s = np.asarray(s, dtype=np.float64).reshape(-1, 21)
    V, R, om = s[:, 3:6], s[:, 6:15].reshape(-1, 3, 3), s[:, 15:18]
    dx = V.copy()                                                         # :471
    dV = ((1.0 - m["vel_damp"]) * V - V) / dt + np.array([0.0, 0.0, -GRAV])   # :475
    w = np.einsum("nij,nj->ni", R, om)                                    # :479 body -> world
    W = np.zeros((len(s), 3, 3))                                          # :481
    W[:, 0, 1], W[:, 0, 2], W[:, 1, 0], W[:, 1, 2], W[:, 2, 0], W[:, 2, 1] = -w[:, 2], w[:, 1], w[:, 2], -w[:, 0], -w[:, 1], w[:, 0]
    dR = (W @ R).reshape(-1, 9)                                           # :484
    inertia = m["inertia"]
    F_omega = (1.0 / inertia) * _cross(-om, inertia * om)                 # :488
    damp = np.clip(m["damp_omega_quadratic"] * om ** 2, 0.0, 1.0)         # :489
    dOm = (1.0 - damp) * F_omega                                          # :490
    return np.concatenate([dx, dV, dR, dOm, np.zeros((len(s), 3))], axis=1)


def affine_G(m, s, dim_mode="3D"):
    """quadrotor.py:501-518.  Returns [N, 21, C], C = 1 / 2 / 4 controls for dim_mode 1D / 2D / 3D."""
    s = np.asarray(s, dtype=np.float64).reshape(-1, 21)
    N = len(s)
    R, om = s[:, 6:15].reshape(-1, 3, 3), s[:, 15:18]
    thrust_sum = np.zeros((3, 4))
    thrust_sum[2, :] = 1.0                                                # :194-195
    dV = (R / m["mass"]) @ (m["thrust_max"] * thrust_sum)                 # :509
    ccw_mx = np.zeros((3, 4))
    ccw_mx[2, :] = m["prop_ccw"]                                          # :183-184
    G_omega = (1.0 / m["inertia"])[:, None] * (m["thrust_max"] * m["prop_cross"].T + m["torque_max"] * ccw_mx)   # :188-191
    damp = np.clip(m["damp_omega_quadratic"] * om ** 2, 0.0, 1.0)         # :515
    dOm = (1.0 - damp)[:, :, None] * G_omega[None]                        # :516
    G = np.concatenate([np.zeros((N, 3, 4)), dV, np.zeros((N, 9, 4)), dOm, np.zeros((N, 3, 4))], axis=1)
    return G @ CONTROL_MX[dim_mode]                                       # :518


# ==================================================================================================
# reward (quadrotor.py:544-638; multi variant quadrotor_multi/quadrotor_multi.py:550-604)
# ==================================================================================================
DEFAULT_REW_COEFF = dict(pos=1.0, effort=0.05, action_change=0.0, crash=1.0, orient=1.0, yaw=0.0, rot=0.0,
                         attitude=0.0, spin=0.1, vel=0.0)                               # quadrotor.py:799-806
DEFAULT_REW_COEFF_MULTI = dict(pos=1.0, pos_offset=0.1, pos_log_weight=1.0, pos_linear_weight=0.1, effort=0.01,
                               action_change=0.0, crash=1.0, orient=1.0, yaw=0.0, rot=0.0, attitude=0.0, spin=0.0,
                               vel=0.0)                                                 # quadrotor_multi.py:810-818


def compute_reward(pos, vel, rot, omega, goal, action, action_prev, crashed, dt, rc, extra_cost=None):
    dist = _norm3(goal - pos)
    if "pos_log_weight" in rc:                                            # quadrotor_multi.py:554
        cost_pos = rc["pos"] * (rc["pos_log_weight"] * np.log(dist + rc["pos_offset"]) + rc["pos_linear_weight"] * dist)
    else:
        cost_pos = rc["pos"] * dist                                       # :547-549
    cost_effort = rc["effort"] * np.sqrt(np.sum(action ** 2, axis=1))     # :553-554
    dact = action - action_prev
    cost_act_change = rc["action_change"] * (dact[:, 0] ** 2 + dact[:, 1] ** 2 + dact[:, 2] ** 2 + dact[:, 3] ** 2) ** 0.5
    cost_vel = rc["vel"] * _norm3(vel)                                    # :562-563
    cost_orient = rc["orient"] * (-rot[:, 2, 2])                          # :567-568
    cost_yaw = rc["yaw"] * (-rot[:, 0, 0])                                # :570-571
    rot_cos = ((rot[:, 0, 0] + rot[:, 1, 1] + rot[:, 2, 2]) - 1.0) / 2.0  # :575
    cost_rotation = rc["rot"] * np.arccos(np.clip(rot_cos, -1.0, 1.0))    # :577-578
    cost_attitude = rc["attitude"] * np.arccos(np.clip(rot[:, 2, 2], -1.0, 1.0))       # :580-581
    cost_spin = rc["spin"] * (omega[:, 0] ** 2 + omega[:, 1] ** 2 + omega[:, 2] ** 2) ** 0.5   # :585-586
    cost_crash = rc["crash"] * crashed.astype(np.float64)                 # :590-591
    terms = [cost_pos, cost_effort, cost_crash, cost_orient, cost_yaw, cost_rotation, cost_attitude, cost_spin,
             cost_act_change, cost_vel]                                   # :593-604 (same summation order)
    if extra_cost is not None:
        terms.append(extra_cost)
    return -dt * np.sum(np.stack(terms, axis=0), axis=0)


# ==================================================================================================
# the env
# ==================================================================================================
class OracleVecEnv(object):
    """N independent copies of the reference QuadrotorEnv (quadrotor.py:647-1156), stepped in lock-step.

    Semantics added on top of the reference (documented in DESIGN.md): auto_reset folds reset() into
    step() - on done the terminal step's reward/done are returned with the post-reset observation.
    n_agents > 1 turns consecutive groups of n_agents envs into one swarm env (collide_swarm).
    """

    def __init__(self, params, n_envs, controller=CTRL_RAW, zero_middle=True, sim_freq=200.0, sim_steps=2,
                 obs_repr="xyz_vxyz_R_omega", ep_time=7, room_size=10.0, init_random_state=False, rew_coeff=None,
                 sense_noise=None, gravity=GRAV, auto_reset=True, goal=(0.0, 0.0, 2.0), box=2.0,
                 multi_reward=False, n_agents=1, collision_k=1.0, collision_coeff=0.0, goals=None, dim_mode="3D",
                 resample_goal=False, excite=False, dynamics_simplification=False, t2w_std=0.005, t2t_std=0.0005,
                 pitch_max=0.5, roll_max=0.5, yaw_max=3.14):
        self.m = model_constants(params, simplified=dynamics_simplification)
        self.m["gravity"] = gravity
        self.N = int(n_envs)
        self.controller = controller
        self.dim_mode = dim_mode
        # init_random_state: False / True (quadrotor.py:1100-1115), or "pitch_roll_restricted" - the law of
        # QuadrotorDynamics.pitch_roll_restricted_random_state (:243-259), which the reference env never selects itself
        self.init_mode = (INIT_PITCH_ROLL_RESTRICTED if init_random_state == "pitch_roll_restricted"
                          else (INIT_RANDOM if init_random_state else INIT_DEFAULT))
        self.t2w_std, self.t2t_std = float(t2w_std), float(t2t_std)       # quadrotor.py:658, :706-712
        self.pitch_max, self.roll_max, self.yaw_max = float(pitch_max), float(roll_max), float(yaw_max)
        self.zero_middle = zero_middle
        self.resample_goal, self.excite = bool(resample_goal), bool(excite)   # quadrotor.py:1078-1081, :957-962
        self.goal_default = np.array(goal, dtype=np.float64)
        self.sim_steps = int(sim_steps)
        self.dt = 1.0 / sim_freq                                          # quadrotor.py:790
        self.ep_len = int(ep_time / (self.dt * self.sim_steps))           # :792
        self.obs_repr = obs_repr
        self.room_box = np.array([[-room_size, -room_size, 0.0], [room_size, room_size, room_size]])   # :723
        self.init_random_state = bool(init_random_state)
        self.max_init_vel = 1.0                                           # :717
        self.max_init_omega = 2 * np.pi                                   # :718
        self.box = box                                                    # :728
        rc = dict(DEFAULT_REW_COEFF_MULTI if multi_reward else DEFAULT_REW_COEFF)
        if rew_coeff:
            rc.update(rew_coeff)
        self.rew_coeff = {k: float(v) for k, v in rc.items()}
        if sense_noise == "default":                                      # :838-849
            sense_noise = dict(DEFAULT_SENSE_NOISE)
        elif isinstance(sense_noise, dict):
            sense_noise = dict(DEFAULT_SENSE_NOISE, **sense_noise)
        self.sn = sense_noise
        self.auto_reset = auto_reset
        self.n_agents = int(n_agents)
        self.collision_k = collision_k
        self.collision_coeff = collision_coeff
        N = self.N
        if goals is None:
            self.goal = np.tile(np.array(goal, dtype=np.float64), (N, 1))  # :1081
        else:
            self.goal = np.array(goals, dtype=np.float64).reshape(N, 3)
        # dynamics state (quadrotor.py:213-224, :438-440) -------------------------------------
        self.pos = np.zeros((N, 3))
        self.vel = np.zeros((N, 3))
        self.rot = np.tile(np.eye(3), (N, 1, 1))
        self.omega = np.zeros((N, 3))
        # set_state stores omega as a float32 *array* (:223); until step1 replaces it (:404) a few reference
        # expressions that combine it with Python scalars are evaluated in float32 (NumPy weak-scalar promotion):
        # kd_a * e_w in the Mellinger controller (quadrotor_control.py:350) and omega ** 2 (:403).
        self.omega_f32 = np.zeros(N, dtype=bool)
        self.rot_damp = np.zeros((N, 4))
        self.cmds_damp = np.zeros((N, 4))
        self.ou = np.zeros((N, 4))                                        # quad_utils.py:191
        self.accelerometer = np.tile(np.array([0.0, 0.0, GRAV]), (N, 1))  # :221
        self.acc = np.zeros((N, 3))
        self.torque = np.zeros((N, 3))
        self.omega_dot = np.zeros((N, 3))
        self.since_last_svd = 0.0                                         # :104 (global phase, never reset)
        self.gyro_bias = np.zeros((N, 3))                                 # sensor_noise.py:93
        self.tick = np.zeros(N, dtype=np.int64)
        self.act_cur = np.zeros((N, 4))                                   # actions[0], :943-944
        self.act_prev = np.zeros((N, 4))                                  # actions[1] (previous step's action)
        self.crashed = np.zeros(N, dtype=bool)
        self.collided = np.zeros(N, dtype=bool)
        self.ep_return = np.zeros(N)
        self.stats = np.zeros(3)                                          # sum ret, sum ret^2, episodes
        self.last_cmds = np.zeros((N, 4))                                 # controller.action (:92, :362)

    # ---------------------------------------------------------------------------------------------
    def substep(self, cmds, ou_z):
        """quadrotor.py:273-436 (step1), one substep for all envs.  ou_z: [N,4] unit normals."""
        m, dt = self.m, self.dt
        c = np.clip(cmds, 0.0, 1.0)                                       # :279
        tau_up = 4 * dt / (m["damp_time_up"] + EPS)                       # :284
        tau_down = 4 * dt / (m["damp_time_down"] + EPS)                   # :285
        tau = np.where(c < self.cmds_damp, tau_down, tau_up)              # :286-287
        tau = np.where(tau > 1.0, 1.0, tau)                               # :288
        thrust_rot = c ** 0.5                                             # :294
        self.rot_damp = tau * (thrust_rot - self.rot_damp) + self.rot_damp  # :295
        cd = self.rot_damp ** 2                                           # :296
        self.ou = self.ou + (m["ou_theta"] * (0.0 - self.ou) + m["ou_sigma"] * ou_z)   # quad_utils.py:197-201
        cd = np.clip(cd + c * self.ou, 0.0, 1.0)                          # :299-300
        self.cmds_damp = cd
        lin = m["linearity"]
        thrusts = m["thrust_max"] * ((1 - lin) * cd ** 2 + lin * cd)      # :302, :140
        torques = m["prop_cross"][None, :, :] * thrusts[:, :, None]       # :304
        torques[:, :, 2] += m["torque_max"] * m["prop_ccw"] * cd          # :307
        thrust_torque = np.sum(torques, axis=1)                           # :310
        N = self.N
        if m["C_drag"] != 0 or m["C_roll"] != 0:                          # :318-349
            vel_body = np.einsum("nji,nj->ni", self.rot, self.vel)        # rot.T @ vel
            pp = m["prop_pos"]
            om = self.omega
            v_rotors = np.empty((N, 4, 3))
            for k in range(4):                                            # quad_utils.py:165-170
                v_rotors[:, k, :] = vel_body + _cross(om, np.tile(pp[k], (N, 1)))
            v_rotors[:, :, 2] = 0.0                                       # :325
            sq = np.sqrt(cd)[:, :, None]
            drag_fi = -m["C_drag"] * sq * v_rotors                        # :328
            drag_force = np.sum(drag_fi, axis=1)                          # :329
            drag_ti = np.empty((N, 4, 3))
            for k in range(4):                                            # :331
                drag_ti[:, k, :] = _cross(drag_fi[:, k, :], np.tile(pp[k], (N, 1)))
            drag_torque = np.sum(drag_ti, axis=1)                         # :332
            roll_torque = np.sum(-m["C_roll"] * m["prop_ccw"][None, :, None] * sq * v_rotors, axis=1)   # :334-335
            visc_torque = drag_torque + roll_torque                       # :336
            vel_norm = np.sqrt(np.sum(vel_body ** 2, axis=1))             # :339
            rdf_norm = np.sqrt(np.sum(drag_force ** 2, axis=1))           # :340
            rdf_clip = np.clip(rdf_norm, 0.0, vel_norm * m["mass"] / (2 * dt))   # :341
            big = rdf_norm > EPS                                          # :342-343
            drag_force = np.where(big[:, None], (drag_force / np.where(big, rdf_norm, 1.0)[:, None]) * rdf_clip[:, None],
                                  drag_force)
            rvt_norm = np.sqrt(np.sum(visc_torque ** 2, axis=1))          # :346
            rvt_clip = np.clip(rvt_norm, 0.0, np.sqrt(np.sum((self.omega * m["inertia"]) ** 2, axis=1)) / (2 * dt))
            big = rvt_norm > EPS                                          # :348-349
            visc_torque = np.where(big[:, None], (visc_torque / np.where(big, rvt_norm, 1.0)[:, None]) * rvt_clip[:, None],
                                   visc_torque)
        else:
            visc_torque = np.zeros((N, 3))                                # :356
            drag_force = np.zeros((N, 3))
        self.torque = thrust_torque + visc_torque                         # :361
        thrust = np.zeros((N, 3))
        thrust[:, 2] = np.sum(thrusts, axis=1)                            # :362
        # rotational dynamics ---------------------------------------------------------------------
        omega_vec = np.einsum("nij,nj->ni", self.rot, self.omega)         # :370
        wx, wy, wz = omega_vec[:, 0], omega_vec[:, 1], omega_vec[:, 2]
        omega_norm = np.sqrt(wx * wx + wy * wy + wz * wz)                 # :372
        nz = omega_norm != 0                                              # :373
        if np.any(nz):
            on = np.where(nz, omega_norm, 1.0)
            K = np.zeros((N, 3, 3))                                       # :375
            K[:, 0, 1] = -wz / on
            K[:, 0, 2] = wy / on
            K[:, 1, 0] = wz / on
            K[:, 1, 2] = -wx / on
            K[:, 2, 0] = -wy / on
            K[:, 2, 1] = wx / on
            ang = omega_norm * dt                                         # :376
            dR = np.eye(3)[None] + np.sin(ang)[:, None, None] * K + (1.0 - np.cos(ang))[:, None, None] * (K @ K)   # :377
            self.rot = np.where(nz[:, None, None], dR @ self.rot, self.rot)   # :378
        self.since_last_svd += dt                                         # :381
        if self.since_last_svd > 0.5:                                     # :382 (since_last_svd_limit :105)
            self.rot = polar_orthonormalize(self.rot)                     # :384-385
            self.since_last_svd = 0                                       # :386
        inertia = m["inertia"]
        self.omega_dot = (1.0 / inertia) * (_cross(-self.omega, inertia * self.omega) + self.torque)   # :398-399
        om2 = np.where(self.omega_f32[:, None], (self.omega.astype(np.float32) ** 2).astype(np.float64), self.omega ** 2)
        odq = np.clip(m["damp_omega_quadratic"] * om2, 0.0, 1.0)          # :403
        self.omega_f32 = np.zeros(N, dtype=bool)
        self.omega = self.omega + (1.0 - odq) * dt * self.omega_dot       # :404
        self.omega = np.clip(self.omega, -m["omega_max"], m["omega_max"])  # :405
        # translational dynamics ------------------------------------------------------------------
        self.pos = self.pos + dt * self.vel                               # :418
        self.pos = np.clip(self.pos, self.room_box[0], self.room_box[1])  # :422
        acc = np.array([0.0, 0.0, -GRAV]) + (1.0 / m["mass"]) * np.einsum("nij,nj->ni", self.rot, thrust + drag_force)  # :426
        self.acc = acc
        self.vel = (1.0 - m["vel_damp"]) * self.vel + dt * acc            # :431
        self.accelerometer = np.einsum("nji,nj->ni", self.rot, acc + np.array([0.0, 0.0, m["gravity"]]))   # :436

    def dynamics_step(self, cmds, ou_z):
        """quadrotor.py:261-262.  ou_z: [N, sim_steps, 4]."""
        for s in range(self.sim_steps):
            self.substep(cmds, ou_z[:, s, :])

    # ---------------------------------------------------------------------------------------------
    def control(self, action):
        """action [N, A] -> normalized motor commands [N,4] (quadrotor_control.py)."""
        m, c = self.m, self.controller
        N = self.N
        if c in (CTRL_RAW, CTRL_VERTICAL, CTRL_VERTPLANE):                # :72-99, :116-136, :159-179
            if self.zero_middle:
                a = 0.5 * (action + 1.0)
                a = np.clip(a, -1.0, 1.0)                                  # low = -1 in zero-middle mode (:80, :90)
            else:
                a = np.clip(1.0 * (action + 0.0), 0.0, 1.0)
            if c == CTRL_VERTICAL:
                a = np.repeat(a[:, 0:1], 4, axis=1)                       # :136
            elif c == CTRL_VERTPLANE:
                a = np.stack([a[:, 0], a[:, 0], a[:, 1], a[:, 1]], axis=1)   # :179
            return a
        if c == CTRL_SHIFTED:                                             # :23-27
            return np.clip((action + 1.0) / m["thrust_to_weight"], 0.0, 1.0)
        if c == CTRL_OMEGA_THRUST:                                        # :225-234
            dw_des = -5.0 * (self.omega - action[:, 1:4])
            acc_des = GRAV * (action[:, 0] + 1.0)
            des = np.concatenate([acc_des[:, None], dw_des], axis=1)
            return np.clip(des @ m["Jinv"].T, 0.0, 1.0)
        if c in (CTRL_VELOCITY_YAW, CTRL_MELLINGER):
            R = self.rot
            if c == CTRL_VELOCITY_YAW:                                    # :250-281 (thrust_mag restated per :275)
                acc_des = -5.0 * (self.vel - action[:, :3]) + np.array([0.0, 0.0, GRAV])
                xc = R[:, :, 0]
                kp_a, kd_a = 100.0, 50.0
                e_w = self.omega - np.concatenate([np.zeros((N, 2)), action[:, 3:4]], axis=1)
                yaw_scale = 1.0
            else:                                                         # :315-362
                to_goal = self.goal - self.pos
                e_p = -_clamp_norm(to_goal, 4.0)                          # :320
                acc_des = -4.5 * e_p - 3.5 * self.vel + np.array([0.0, 0.0, GRAV])    # :323
                xc = np.tile(np.array([1.0, 0.0, 0.0]), (N, 1))           # :333 (rot_des = eye)
                kp_a, kd_a = 200.0, 50.0
                e_w = self.omega
                yaw_scale = 0.2                                           # :347
            zb = _normalize(acc_des)                                      # :338
            yb = _normalize(_cross(zb, xc))                               # :339
            xb = _cross(yb, zb)                                           # :340
            R_des = np.stack([xb, yb, zb], axis=2)                        # :341 column_stack
            M = np.einsum("nji,njk->nik", R_des, R) - np.einsum("nji,njk->nik", R, R_des)   # :346
            e_R = 0.5 * np.stack([M[:, 2, 1], M[:, 0, 2], M[:, 1, 0]], axis=1)
            e_R[:, 2] *= yaw_scale
            kd_ew = kd_a * e_w
            if c == CTRL_MELLINGER:                                       # float32 product while omega is the f32 array
                kd_ew = np.where(self.omega_f32[:, None],
                                 (np.float32(kd_a) * e_w.astype(np.float32)).astype(np.float64), kd_ew)
            dw_des = -kp_a * e_R - kd_ew                                  # :350
            thrust_mag = np.sum(acc_des * R[:, :, 2], axis=1)             # :352 / :275
            des = np.concatenate([thrust_mag[:, None], dw_des], axis=1)
            return np.clip(des @ m["Jinv"].T, 0.0, 1.0)                   # :357-359
        raise ValueError(c)

    # ---------------------------------------------------------------------------------------------
    def observe(self, obs_z, n_calls):
        rel, h, acc_, act, quat, t2w, t2t = obs_layout(self.obs_repr)
        rot_in = R2quat(self.rot) if quat else self.rot                   # get_state.py:277 (quat builders)
        pos, vel, rot, omega, acc, gyro_bias = add_noise(
            self.sn, self.pos, self.vel, rot_in, self.omega, self.accelerometer, self.dt, obs_z, self.gyro_bias,
            n_calls=n_calls)
        e = pos - self.goal
        v = vel
        if rel:                                                           # get_state.py:159-160 (true R)
            e = np.einsum("nji,nj->ni", self.rot, e)
            v = np.einsum("nji,nj->ni", self.rot, vel)
        parts = [e, v, rot.reshape(self.N, -1), omega]
        if h:
            parts.append(pos[:, 2:3])                                     # get_state.py:143 (noisy z)
        if acc_:
            parts.append(acc)
        if act:
            parts.append(self.act_prev)                                   # get_state.py:246,259 -> actions[1]
        if t2w:                                                           # get_state.py:334-337: drawn whether or not the
            w = self.m["thrust_to_weight"]                                # SensorNoise is bypassed
            nw = np.clip(w + abs((self.t2w_std / 2) * w) * obs_z[:, Z_NT2W:Z_NT2W + 1], T2W_MIN, T2W_MAX)
            parts.append((nw - T2W_MIN) / (T2W_MAX - T2W_MIN))
        if t2t:                                                           # get_state.py:379-382
            w = self.m["torque_to_thrust"]
            nw = np.clip(w + abs((self.t2t_std / 2) * w) * obs_z[:, Z_NT2T:Z_NT2T + 1], T2T_MIN, T2T_MAX)
            parts.append((nw - T2T_MIN) / (T2T_MAX - T2T_MIN))
        return np.concatenate(parts, axis=1), gyro_bias

    # ---------------------------------------------------------------------------------------------
    def reset(self, mask=None, reset_z=None, goal_z=None):
        """quadrotor.py:1059-1144 for the envs in mask.  reset_z: [N, RESET_Z] unit draws; returns obs [N,O]."""
        N = self.N
        if mask is None:
            mask = np.ones(N, dtype=bool)
        if reset_z is None:
            reset_z = np.zeros((N, RESET_Z))
        if self.resample_goal or self.excite:                             # :1078-1081: the goal is re-made at every reset
            g = np.tile(self.goal_default, (N, 1))
            if self.resample_goal:                                        # np.random.uniform(0.5, 2.0)
                u = np.zeros(N) if goal_z is None else np.asarray(goal_z, dtype=np.float64).reshape(N, GOAL_Z)[:, 3]
                g = np.stack([np.zeros(N), np.zeros(N), 1.25 + 0.75 * u], axis=1)
            self.goal = np.where(mask[:, None], g, self.goal)
        goal = self.goal
        pos = self.box * reset_z[:, RZ_UPOS:RZ_UPOS + 3] + goal           # :1087
        if self.dim_mode == "1D":                                         # :1089-1092
            pos[:, 0:2] = goal[:, 0:2]
        elif self.dim_mode == "2D":
            pos[:, 1] = goal[:, 1]
        pos[:, 2] = np.where(pos[:, 2] < 0.25, 0.25, pos[:, 2])           # :1094
        u01 = lambda k: 0.5 + 0.5 * reset_z[:, k]                         # np.random.rand() from a unit symmetric draw
        if self.init_mode != INIT_DEFAULT and self.dim_mode == "1D":      # :1101-1103
            omega = np.zeros((N, 3))
            rot = np.tile(np.eye(3), (N, 1, 1))
            vel = np.stack([np.zeros(N), np.zeros(N), self.max_init_vel * u01(RZ_UVELMAG)], axis=1)
        elif self.init_mode != INIT_DEFAULT and self.dim_mode == "2D":    # :1104-1110 (draw order: omega, vel, theta)
            omega = np.stack([np.zeros(N), self.max_init_omega * u01(RZ_UOMMAG), np.zeros(N)], axis=1)
            vel = self.max_init_vel * np.stack([u01(RZ_UVEL), np.zeros(N), u01(RZ_UVEL + 2)], axis=1)
            theta = np.pi * u01(RZ_YAW)
            c, s_ = np.cos(theta), np.sin(theta)
            rot = np.zeros((N, 3, 3))
            rot[:, 0, 0], rot[:, 0, 2], rot[:, 1, 1], rot[:, 2, 0], rot[:, 2, 2] = c, -s_, 1.0, s_, c
        elif self.init_mode != INIT_DEFAULT:                              # :1113-1115 -> :227-239 / :243-259
            vm, om_ = self.max_init_vel, self.max_init_omega
            vel = vm * reset_z[:, RZ_UVEL:RZ_UVEL + 3]                    # :230
            vel_magn = vm * (0.5 + 0.5 * reset_z[:, RZ_UVELMAG])          # :231 uniform(0, vel_max)
            vel = (vel_magn / (np.sqrt(np.sum(vel ** 2, axis=1)) + EPS))[:, None] * vel   # :232
            omega = om_ * reset_z[:, RZ_UOM:RZ_UOM + 3]                   # :234
            om_magn = om_ * (0.5 + 0.5 * reset_z[:, RZ_UOMMAG])           # :235
            omega = (om_magn / (np.sqrt(np.sum(omega ** 2, axis=1)) + EPS))[:, None] * omega   # :236
            if self.init_mode == INIT_PITCH_ROLL_RESTRICTED:              # :254-257 (draw order: pitch, roll, yaw)
                rot = euler2mat_sxyz(self.roll_max * reset_z[:, RZ_NUP + 1], self.pitch_max * reset_z[:, RZ_NUP],
                                     self.yaw_max * reset_z[:, RZ_NUP + 2])
            else:
                up = _normalize(reset_z[:, RZ_NUP:RZ_NUP + 3])            # quad_utils.py:47-57 (accepted draws)
                fwd = _normalize(reset_z[:, RZ_NFWD:RZ_NFWD + 3])
                left = _normalize(_cross(up, fwd))
                up = _cross(fwd, left)
                rot = np.stack([fwd, left, up], axis=2)
        else:                                                             # :1118-1126
            vel = np.zeros((N, 3))
            omega = np.zeros((N, 3))
            psi = np.pi * reset_z[:, RZ_YAW]                              # accepted randyaw draw, quad_utils.py:139-141
            if self.dim_mode != "3D":                                     # :1120-1121 rotation = eye
                psi = np.zeros(N)
            cpsi, spsi = np.cos(psi), np.sin(psi)
            rot = np.zeros((N, 3, 3))
            rot[:, 0, 0] = cpsi
            rot[:, 0, 1] = -spsi
            rot[:, 1, 0] = spsi
            rot[:, 1, 1] = cpsi
            rot[:, 2, 2] = 1.0
        omega = omega.astype(np.float32).astype(np.float64)               # :223 float32 cast in set_state
        mk = mask
        self.pos = np.where(mk[:, None], pos, self.pos)
        self.vel = np.where(mk[:, None], vel, self.vel)
        self.rot = np.where(mk[:, None, None], rot, self.rot)
        self.omega = np.where(mk[:, None], omega, self.omega)
        self.omega_f32 = np.where(mk, True, self.omega_f32)
        self.acc = np.where(mk[:, None], 0.0, self.acc)                   # :220
        self.accelerometer = np.where(mk[:, None], np.array([0.0, 0.0, GRAV]), self.accelerometer)   # :221
        self.rot_damp = np.where(mk[:, None], 0.0, self.rot_damp)         # :438-440
        self.cmds_damp = np.where(mk[:, None], 0.0, self.cmds_damp)
        self.crashed = np.where(mk, False, self.crashed)                  # :1139
        self.tick = np.where(mk, 0, self.tick)                            # :1140
        self.act_cur = np.where(mk[:, None], 0.0, self.act_cur)   # :1141
        self.act_prev = np.where(mk[:, None], 0.0, self.act_prev)
        self.ep_return = np.where(mk, 0.0, self.ep_return)
        obs, gb = self.observe(reset_z[:, RZ_OBS:RZ_OBS + OBS_Z], n_calls=1)   # :1143
        self.gyro_bias = np.where(mk[:, None], gb, self.gyro_bias)
        return obs

    def set_state(self, pos, vel, rot, omega):
        """quadrotor.py:213-224 + :438-440 for all envs (test hook for known-answer states)."""
        self.pos = np.array(pos, dtype=np.float64).reshape(self.N, 3)
        self.vel = np.array(vel, dtype=np.float64).reshape(self.N, 3)
        self.rot = np.array(rot, dtype=np.float64).reshape(self.N, 3, 3)
        self.omega = np.array(omega, dtype=np.float64).reshape(self.N, 3).astype(np.float32).astype(np.float64)
        self.omega_f32 = np.ones(self.N, dtype=bool)
        self.acc = np.zeros((self.N, 3))
        self.accelerometer = np.tile(np.array([0.0, 0.0, GRAV]), (self.N, 1))
        self.rot_damp = np.zeros((self.N, 4))
        self.cmds_damp = np.zeros((self.N, 4))

    # ---------------------------------------------------------------------------------------------
    def collide_swarm(self):
        """Builder-authored spec (no reference counterpart).

        Within each group of n_agents consecutive envs: drones i != j collide when
        d_ij = |p_i - p_j| < collision_k * 2 * arm.  Response: equal-mass elastic exchange of the
        normal velocity component, evaluated from the PRE-collision velocities of all drones
        (order-independent): n = (p_i - p_j) / max(d, 1e-5), delta = (v_j - v_i) . n,
        v_i += delta * n summed over every j that i collides with and approaches (delta > 0).
        Sets self.collided (per drone, this step).
        """
        M = self.n_agents
        if M <= 1:
            self.collided = np.zeros(self.N, dtype=bool)
            return
        E = self.N // M
        p = self.pos.reshape(E, M, 3)
        v = self.vel.reshape(E, M, 3)
        # one model per oracle env: every pair touches at collision_k * 2 * arm.  arm_env ([N], optional): drones with their
        # own models (per-env dynamics) - the pair (i, j) touches at collision_k * (arm_i + arm_j)
        arm = getattr(self, "arm_env", None)
        arm = np.full((E, M), self.m["arm"]) if arm is None else np.asarray(arm, dtype=np.float64).reshape(E, M)
        dv = np.zeros_like(v)
        hit = np.zeros((E, M), dtype=bool)
        for i in range(M):
            for j in range(M):
                if i == j:
                    continue
                d = p[:, i] - p[:, j]
                dist = np.sqrt(np.sum(d ** 2, axis=1))
                n = d / np.maximum(dist, 1e-5)[:, None]
                delta = np.sum((v[:, j] - v[:, i]) * n, axis=1)
                c = dist < self.collision_k * arm[:, i] + self.collision_k * arm[:, j]
                hit[:, i] |= c
                dv[:, i] += np.where((c & (delta > 0))[:, None], delta[:, None] * n, 0.0)
        self.vel = (v + dv).reshape(self.N, 3)
        self.collided = hit.reshape(self.N)

    # ---------------------------------------------------------------------------------------------
    def step(self, action, ou_z=None, obs_z=None, reset_z=None, goal_z=None):
        """quadrotor.py:942-1028 (+ auto-reset).  Returns obs [N,O], reward [N], done [N] (bool)."""
        N = self.N
        action = np.asarray(action, dtype=np.float64).reshape(N, -1)
        if ou_z is None:
            ou_z = np.zeros((N, self.sim_steps, 4))
        if obs_z is None:
            obs_z = np.zeros((N, OBS_Z))
        act4 = np.zeros((N, 4))
        act4[:, :action.shape[1]] = action
        self.act_prev = self.act_cur.copy()                   # :943 actions[1] = actions[0]
        self.act_cur = act4.copy()                                    # :944 actions[0] = action
        if self.excite:                                                   # :957-962, every 5th tick (before the increment)
            gz = np.zeros((N, GOAL_Z)) if goal_z is None else np.asarray(goal_z, dtype=np.float64).reshape(N, GOAL_Z)
            new = np.stack([0.5 * gz[:, 0], 0.5 * gz[:, 1], 2.0 + 0.5 * gz[:, 2]], axis=1)
            self.goal = np.where((self.tick % 5 == 0)[:, None], new, self.goal)
        cmds = self.control(action)                                       # :966-970
        self.last_cmds = cmds
        self.dynamics_step(cmds, ou_z)
        self.collide_swarm()
        self.crashed = self.pos[:, 2] <= self.m["arm"]                    # :977 (the :978-981 clause is dead)
        extra = None
        if self.n_agents > 1 and self.collision_coeff != 0.0:
            extra = self.collision_coeff * self.collided.astype(np.float64)
        reward = compute_reward(self.pos, self.vel, self.rot, self.omega, self.goal, act4, self.act_prev,
                                self.crashed, self.dt, self.rew_coeff, extra_cost=extra)   # :984
        self.tick = self.tick + 1                                         # :986
        done = self.tick > self.ep_len                                    # :987
        self.ep_return = self.ep_return + reward
        obs, self.gyro_bias = self.observe(obs_z, n_calls=3)              # :947, :970, :988
        if np.any(done):
            r = self.ep_return[done]
            self.stats += np.array([np.sum(r), np.sum(r * r), float(r.size)])
            if self.auto_reset:
                obs_reset = self.reset(mask=done, reset_z=reset_z, goal_z=goal_z)
                obs = np.where(done[:, None], obs_reset, obs)
        return obs, reward, done

"""Derived dynamics constants: parameter dict -> flat record -> QsModel (host side, runs once per model).

Follows QuadrotorDynamics.__init__/update_model (gym_art/quadrotor/quadrotor.py:77-208) and quadrotor_jacobian
(quadrotor_control.py:192-203); see SURVEY.md section 8 row a0 for the known answers the tests pin.
"""
import numpy as np

from .. import _lib
from .inertia import QuadLink, QuadLinkSimplified

GRAV = 9.81
EPS = 1e-6
PROP_CCW = np.array([-1., 1., -1., 1.])        # quadrotor.py:88
OMEGA_MAX = 40.                                # quadrotor.py:91
OU_THETA = 0.15                                # quad_utils.py:181


class ModelConstants(object):
    """Everything the kernels need to know about one quadrotor model."""

    def __init__(self, params, gravity=GRAV, dynamics_simplification=False):
        self.params = params
        # quadrotor.py:143-146
        self.link = QuadLinkSimplified(params=params["geom"]) if dynamics_simplification else QuadLink(params=params["geom"])
        motor, damp, noise = params["motor"], params["damp"], params["noise"]
        self.mass = float(self.link.m)
        self.inertia = np.array(np.diagonal(self.link.I_com), dtype=np.float64)
        self.thrust_to_weight = float(motor["thrust_to_weight"])
        self.torque_to_thrust = float(motor["torque_to_thrust"])
        self.motor_linearity = float(motor["linearity"])
        self.C_rot_drag = float(motor["C_drag"])
        self.C_rot_roll = float(motor["C_roll"])
        self.motor_damp_time_up = float(motor["damp_time_up"])
        self.motor_damp_time_down = float(motor["damp_time_down"])
        self.thrust_noise_ratio = float(noise["thrust_noise_ratio"])
        self.vel_damp = float(damp["vel"])
        self.damp_omega_quadratic = float(damp["omega_quadratic"])
        asym = np.array(motor.get("assymetry", [1.0, 1.0, 1.0, 1.0]), dtype=np.float64)
        self.motor_assymetry = asym * 4. / np.sum(asym)                                   # :174
        self.thrust_max = GRAV * self.mass * self.thrust_to_weight * self.motor_assymetry / 4.0   # :175
        self.torque_max = self.torque_to_thrust * self.thrust_max                        # :176
        self.prop_pos = np.array(self.link.prop_pos, dtype=np.float64)
        self.prop_ccw = PROP_CCW.copy()
        self.prop_crossproducts = np.cross(self.prop_pos, [0., 0., 1.])                  # :182
        self.ou_sigma = 0.2 * self.thrust_noise_ratio                                    # :198
        self.ou_theta = OU_THETA
        self.arm = float(np.linalg.norm(self.link.motor_xyz[:2]))                        # :200
        self.omega_max = OMEGA_MAX
        self.gravity = float(gravity)
        # jacobian of (thrust acceleration, angular acceleration) w.r.t. normalized motor thrusts
        torque = self.thrust_max * self.prop_crossproducts.T
        torque[2, :] = self.torque_max * self.prop_ccw
        self.jacobian = np.vstack([self.thrust_max * np.ones((1, 4)) / self.mass, (1.0 / self.inertia)[:, None] * torque])
        self.jacobian_cond = float(np.linalg.cond(self.jacobian))
        self.Jinv = np.linalg.inv(self.jacobian)
        # forced-dynamics helpers exposed by the reference object (quadrotor.py:186-205)
        ccw_mx = np.zeros((3, 4))
        ccw_mx[2, :] = self.prop_ccw
        self.G_omega = (1.0 / self.inertia)[:, None] * (self.thrust_max * self.prop_crossproducts.T + self.torque_max * ccw_mx)
        self.torque_to_inertia = np.sum(self.G_omega @ np.array([[0, 0, 0], [0, 1, 1], [1, 1, 0], [1, 0, 1]]), axis=1)

    def rows(self, dt):
        """This model as one column of the per-env model buffer: float64 [QS_MODEL_ROWS, 4] (layout in quadsim.h)."""
        r = np.zeros((_lib.QS_MODEL_ROWS, 4))
        r[0] = self.thrust_max
        r[1] = self.torque_max * self.prop_ccw
        r[2] = self.prop_pos[:, 1]
        r[3] = -self.prop_pos[:, 0]
        r[4] = self.prop_pos[:, 2]
        r[5, :3], r[5, 3] = self.inertia, self.mass
        r[6, :3], r[6, 3] = 1.0 / self.inertia, 1.0 / self.mass
        r[7] = [min(1.0, 4.0 * dt / (self.motor_damp_time_up + EPS)), min(1.0, 4.0 * dt / (self.motor_damp_time_down + EPS)),
                self.ou_sigma, self.arm]                                                   # quadrotor.py:284-288
        r[8] = [self.motor_linearity, self.C_rot_drag, self.C_rot_roll, self.thrust_to_weight]
        r[9, :3] = [self.vel_damp, self.damp_omega_quadratic, self.torque_to_thrust]
        r[10:14] = self.Jinv
        return r

    def to_qs(self):
        m = _lib.QsModel()
        m.mass = self.mass
        for k in range(3):
            m.inertia[k] = self.inertia[k]
        for k in range(4):
            m.thrust_max[k] = self.thrust_max[k]
            m.torque_max[k] = self.torque_max[k]
            m.prop_ccw[k] = self.prop_ccw[k]
            for j in range(3):
                m.prop_pos[k][j] = self.prop_pos[k, j]
            for j in range(4):
                m.Jinv[k][j] = self.Jinv[k, j]
        m.linearity, m.C_drag, m.C_roll = self.motor_linearity, self.C_rot_drag, self.C_rot_roll
        m.damp_time_up, m.damp_time_down = self.motor_damp_time_up, self.motor_damp_time_down
        m.vel_damp, m.damp_omega_quadratic = self.vel_damp, self.damp_omega_quadratic
        m.ou_theta, m.ou_sigma = self.ou_theta, self.ou_sigma
        m.arm, m.omega_max, m.gravity = self.arm, self.omega_max, self.gravity
        m.thrust_to_weight = self.thrust_to_weight
        m.torque_to_thrust = self.torque_to_thrust
        return m

"""Controllers with the reference's signatures (gym_art/quadrotor/quadrotor_control.py), computed on the B200.

Each class keeps the constructor, `action_space(dynamics)` and `step`/`step_func` call shape of its reference
counterpart; the mixing arithmetic itself (action -> 4 normalized motor commands) runs inside the CUDA kernels
(csrc/qs_device.cuh controller_cmds / attitude_controller), followed - exactly as in the reference, where every
controller ends with `dynamics.step(thrusts, dt)` - by `dynamics_steps_num` substeps of the dynamics.
`dynamics` is a gym_art_b200.quadrotor.quadrotor.QuadrotorDynamics (one drone or a batch).
"""
import numpy as np

from .spaces import Box

GRAV = 9.81


def quadrotor_jacobian(dynamics):
    """quadrotor_control.py:192-203 - d(thrust acceleration, angular acceleration) / d(normalized motor thrusts)."""
    J = np.array(dynamics.model.jacobian, dtype=np.float64)
    if np.linalg.cond(J) > 50:                                             # :199-202 warns, does not fail
        print("WARN: Jacobian conditioning is high: ", np.linalg.cond(J))
    return J


class _DeviceController(object):
    """Shared plumbing: run `name`'s mixing + the dynamics on the device and remember the motor commands."""
    name = None
    zero_middle = True

    def _run(self, dynamics, action, dt, goal=None):
        self.action = dynamics._controller_step(self.name, action, dt, goal=goal, zero_middle=self.zero_middle)


class ShiftedMotorControl(_DeviceController):                               # :11-27
    name = "shifted"

    def __init__(self, dynamics):
        self.action = None

    def action_space(self, dynamics):
        low = -1.0 * np.ones(4)
        high = (dynamics.thrust_to_weight - 1.0) * np.ones(4)
        return Box(low, high, dtype=np.float32)

    def step(self, dynamics, action, dt):
        self._run(dynamics, action, dt)


class RawControl(_DeviceController):                                        # :65-99 (the second definition wins)
    name = "raw"

    def __init__(self, dynamics, zero_action_middle=True, dim_mode="3D"):
        self.zero_action_middle = zero_action_middle
        self.zero_middle = bool(zero_action_middle)
        self.action = None
        self.step_func = self.step

    def action_space(self, dynamics):
        if not self.zero_action_middle:
            self.low, self.bias, self.scale = np.zeros(4), 0, 1.0
        else:
            self.low, self.bias, self.scale = -np.ones(4), 1.0, 0.5
        self.high = np.ones(4)
        return Box(self.low, self.high, dtype=np.float32)

    def step(self, dynamics, action, goal, dt, observation=None):
        self._run(dynamics, action, dt)

    step_tf = step


class VerticalControl(_DeviceController):                                   # :102-144
    name = "vertical"

    def __init__(self, dynamics, zero_action_middle=True, dim_mode="3D"):
        self.zero_action_middle = zero_action_middle
        self.zero_middle = bool(zero_action_middle)
        self.dim_mode = dim_mode
        self.action = None
        self.step_func = self.step

    def action_space(self, dynamics):
        if not self.zero_action_middle:
            self.low, self.bias, self.scale = np.zeros(1), 0, 1.0
        else:
            self.low, self.bias, self.scale = -np.ones(1), 1.0, 0.5
        self.high = np.ones(1)
        return Box(self.low, self.high, dtype=np.float32)

    def step(self, dynamics, action, goal, dt, observation=None):
        self._run(dynamics, action, dt)

    step_tf = step


class VertPlaneControl(VerticalControl):                                    # :146-187
    name = "vertplane"

    def action_space(self, dynamics):
        if not self.zero_action_middle:
            self.low, self.bias, self.scale = np.zeros(2), 0, 1.0
        else:
            self.low, self.bias, self.scale = -np.ones(2), 1.0, 0.5
        self.high = np.ones(2)
        return Box(self.low, self.high, dtype=np.float32)


def _rate_thrust_box(dynamics):                                             # :213-221, :514-522
    circle_per_sec = 2 * np.pi
    max_rp, max_yaw = 5 * circle_per_sec, 1 * circle_per_sec
    low = np.array([-1.0, -max_rp, -max_rp, -max_yaw])
    high = np.array([dynamics.thrust_to_weight - 1.0, max_rp, max_rp, max_yaw])
    return Box(low, high, dtype=np.float32)


class OmegaThrustControl(_DeviceController):                                # :208-234
    name = "omega_thrust"

    def __init__(self, dynamics):
        self.Jinv = np.linalg.inv(quadrotor_jacobian(dynamics))
        self.action = None

    def action_space(self, dynamics):
        return _rate_thrust_box(dynamics)

    def step(self, dynamics, action, dt):
        self._run(dynamics, action, dt)


class VelocityYawControl(_DeviceController):                                # :238-281 (thrust_mag as the commented :275)
    name = "velocity_yaw"

    def __init__(self, dynamics):
        self.Jinv = np.linalg.inv(quadrotor_jacobian(dynamics))
        self.action = None

    def action_space(self, dynamics):
        vmax, dymax = 20.0, 4 * np.pi
        high = np.array([vmax, vmax, vmax, dymax])
        return Box(-high, high, dtype=np.float32)

    def step(self, dynamics, action, dt):
        self._run(dynamics, action, dt)


class NonlinearPositionController(_DeviceController):                       # Mellinger, :286-362
    name = "mellinger"

    def __init__(self, dynamics, tf_control=True):
        # the reference's TensorFlow-1 graph variant (:364-511) computes the same law; there is one device path here
        self.Jinv = np.linalg.inv(quadrotor_jacobian(dynamics))
        self.action = None
        self.kp_p, self.kd_p = 4.5, 3.5
        self.kp_a, self.kd_a = 200.0, 50.0
        self.rot_des = np.eye(3)
        self.tf_control = False
        self.step_func = self.step

    def step(self, dynamics, goal, dt, action=None, observation=None):
        self._run(dynamics, None, dt, goal=goal)

    step_tf = step

    def action_space(self, dynamics):
        return _rate_thrust_box(dynamics)

"""Shipped quadrotor parameter sets (host side).

Same nested-dict schema and values as the reference's gym_art/quadrotor/quad_models.py (crazyflie_params :1,
defaultquad_params :45, mediumquad_params :88, crazyflie_lowinertia_params :135): geom / damp / noise / motor.
The numbers are data (Crazyflie 2.0, AscTec Hummingbird-like, a mid-size quad); the table form is ours.
"""
import copy


def _box(l, w, h, m):
    return {"l": l, "w": w, "h": h, "m": m}


def _cyl(h, r, m):
    return {"h": h, "r": r, "m": m}


#            body                         payload                         arms                              motors                     propellers                    motor xy   payload side
_GEOM = {
    "crazyflie": (_box(0.03, 0.03, 0.004, 0.005), _box(0.035, 0.02, 0.008, 0.01), _box(0.022, 0.005, 0.005, 0.001),
                  _cyl(0.02, 0.0035, 0.0015), _cyl(0.002, 0.022, 0.00075), 0.065 / 2, 1),
    "defaultquad": (_box(0.1, 0.1, 0.085, 0.5), _box(0.12, 0.12, 0.04, 0.1), _box(0.1, 0.015, 0.015, 0.025),
                    _cyl(0.02, 0.025, 0.02), _cyl(0.001, 0.1, 0.009), 0.12, -1),
    "mediumquad": (_box(0.04, 0.04, 0.04, 0.04), _box(0.06, 0.015, 0.015, 0.029), _box(0.04, 0.01, 0.003, 0.006),
                   _cyl(0.013, 0.007, 0.006), _cyl(0.007, 0.035, 0.0012), 0.046, -1),
    "crazyflie_lowinertia": (_box(0.03, 0.03, 0.004, 0.014), _box(0.035, 0.02, 0.008, 0.01),
                             _box(0.022, 0.005, 0.005, 0.0005), _cyl(0.02, 0.0035, 0.0005),
                             _cyl(0.002, 0.022, 0.0000075), 0.065 / 2, 1),
}

#            thrust_to_weight, torque_to_thrust, damp_time_up, damp_time_down
_MOTOR = {
    "crazyflie": (1.9, 0.006, 0.15, 0.15),
    "defaultquad": (2.8, 0.05, 0, 0),
    "mediumquad": (2.5, 0.05, 0.15, 0.15),
    "crazyflie_lowinertia": (1.9, 0.006, 0.15, 0.15),
}


def _params(name):
    body, payload, arms, motors, props, mxy, side = copy.deepcopy(_GEOM[name])
    t2w, t2t, up, down = _MOTOR[name]
    geom = {"body": body, "payload": payload, "arms": arms, "motors": motors, "propellers": props,
            "motor_pos": {"xyz": [mxy, mxy, 0.]}, "arms_pos": {"angle": 45., "z": 0.},
            "payload_pos": {"xy": [0., 0.], "z_sign": side}}
    motor = {"thrust_to_weight": t2w, "assymetry": [1.0, 1.0, 1.0, 1.0], "torque_to_thrust": t2t, "linearity": 1.0,
             "C_drag": 0., "C_roll": 0., "damp_time_up": up, "damp_time_down": down}
    return {"geom": geom, "damp": {"vel": 0.0, "omega_quadratic": 0.0}, "noise": {"thrust_noise_ratio": 0.05},
            "motor": motor}


def crazyflie_params():
    return _params("crazyflie")


def defaultquad_params():
    return _params("defaultquad")


def mediumquad_params():
    return _params("mediumquad")


def crazyflie_lowinertia_params():
    return _params("crazyflie_lowinertia")

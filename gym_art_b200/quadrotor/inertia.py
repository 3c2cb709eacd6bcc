"""Mass / centre of mass / inertia / propeller positions of the quadrotor link model (host side, runs once).

Mirrors the *results* of the reference's gym_art/quadrotor/inertia.py QuadLink (:182-310): a central body box, a
payload box, four arm boxes yawed by +-arm_angle, four motor cylinders and four propeller cylinders in X
configuration, motors ordered front-right, back-right, back-left, front-left (:236-241).  Implemented as one
vectorised pass over a link table instead of per-link classes.  Only the diagonal of I_com is consumed by the
dynamics (quadrotor.py:152); the off-diagonals follow the reference's parallel-axis convention, including its
use of I_xy in the xz term (inertia.py:34), so that I_com compares equal to the reference entry by entry.
"""
import numpy as np

MOTOR_SIGNS = np.array([[1., -1., -1., 1.],      # x of FR, BR, BL, FL
                        [-1., -1., 1., 1.],      # y
                        [1., 1., 1., 1.]])


def box_inertia(l, w, h, m):
    """Solid box, l along x, w along y, h along z."""
    return np.array([m * (h ** 2 + w ** 2), m * (l ** 2 + h ** 2), m * (w ** 2 + l ** 2)]) / 12.0


def cylinder_inertia(h, r, m):
    """Solid vertical cylinder."""
    side = m * (3 * r ** 2 + h ** 2) / 12.0
    return np.array([side, side, 0.5 * m * r ** 2])


def _shift(I, m, xyz):
    x, y, z = xyz
    out = np.diag([I[0, 0] + m * (y * y + z * z), I[1, 1] + m * (x * x + z * z), I[2, 2] + m * (x * x + y * y)])
    out[0, 1] = out[1, 0] = I[0, 1] + m * x * y
    out[0, 2] = out[2, 0] = I[0, 1] + m * x * z      # reference convention (inertia.py:34)
    out[1, 2] = out[2, 1] = I[1, 2] + m * y * z
    return out


class QuadLink(object):
    """Attributes: m, com, I_com [3,3], prop_pos [4,3], motor_xyz, arm_length, arm_angle, params."""

    motors_num = 4

    def __init__(self, params=None, verbose=False):
        if params is None:
            from .quad_models import crazyflie_params
            params = crazyflie_params()["geom"]
        self.params = params
        body, payload, arms = params["body"], params["payload"], params["arms"]
        motors, props = params["motors"], params["propellers"]
        angle = float(params["arms_pos"]["angle"]) / 180. * np.pi
        self.arm_angle = angle if angle != 0. else 0.01
        self.motor_xyz = np.array(params["motor_pos"]["xyz"], dtype=np.float64)
        dy = self.motor_xyz[1] - body["w"] / 2.
        if "l" not in arms:
            arms["l"] = dy / np.sin(self.arm_angle)
        self.arm_length = arms["l"]
        arm_xyz = np.array([self.motor_xyz[0] - dy / (2 * np.tan(self.arm_angle)), self.motor_xyz[1] - dy / 2,
                            params["arms_pos"]["z"]])
        motors_xyz = (MOTOR_SIGNS * self.motor_xyz[:, None]).T                      # [4,3]
        props_xyz = motors_xyz.copy()
        props_xyz[:, 2] = props_xyz[:, 2] + motors["h"] / 2. + props["h"]
        arms_xyz = (MOTOR_SIGNS * arm_xyz[:, None]).T
        arm_yaw = np.array([-1., 1., -1., 1.]) * self.arm_angle
        payload_z = np.sign(params["payload_pos"]["z_sign"]) * (body["h"] + payload["h"]) / 2

        # a link gives its mass directly ("m") or through its material density (inertia.py:75-82, :98-99, :160-161):
        # RandomQuad samples densities (quadrotor_randomization.py:155-164)
        def box(d):
            m = d["m"] if d.get("m") is not None else d["density"] * d["l"] * d["w"] * d["h"]
            return m, box_inertia(d["l"], d["w"], d["h"], m)

        def cyl(d):
            m = d["m"] if d.get("m") is not None else d["density"] * np.pi * d["h"] * d["r"] ** 2
            return m, cylinder_inertia(d["h"], d["r"], m)

        parts = [box(body), box(payload)] + [box(arms)] * 4 + [cyl(motors)] * 4 + [cyl(props)] * 4
        # link table: mass, principal inertia, yaw, position
        mass = [m for m, _ in parts]
        diag = [I for _, I in parts]
        yaw = [0., 0.] + list(arm_yaw) + [0.] * 8
        pos = [np.zeros(3), np.array(list(params["payload_pos"]["xy"]) + [payload_z], dtype=np.float64)] + \
              list(arms_xyz) + list(motors_xyz) + list(props_xyz)
        mass = np.array(mass, dtype=np.float64)
        pos = np.array(pos, dtype=np.float64)

        self.m = np.sum(mass)
        self.com = sum(mass[i] * pos[i] for i in range(len(mass))) / self.m
        pos = pos - self.com
        total = np.zeros((3, 3))
        self.links_I = []
        for i in range(len(mass)):
            c, s = np.cos(yaw[i]), np.sin(yaw[i])
            Rz = np.array([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]]) if yaw[i] else np.eye(3)
            I = _shift(Rz @ np.diag(diag[i]) @ Rz.T, mass[i], pos[i])
            self.links_I.append(I)
            total = total + I
        self.I_com = total
        self.prop_pos = pos[6:10].copy()        # the motor centres, as in the reference (inertia.py:309)
        if verbose:
            print("QuadLink: m=%g com=%s diag(I)=%s" % (self.m, self.com, np.diagonal(self.I_com)))


class QuadLinkSimplified(object):
    """The reference's simplified body (inertia.py:312-420, `dynamics_simplification=True`): two perpendicular rods
    carrying the whole mass and four massless propellers.  Attributes as QuadLink: m, com, I_com, prop_pos, motor_xyz."""

    motors_num = 4
    rods_num = 2

    def __init__(self, params=None, verbose=False):
        from .quad_models import crazyflie_params
        base = crazyflie_params()["geom"]                   # the reference fills unspecified parts from CrazyFlie values
        base.update(params or {})
        p = self.params = base
        self.arm_length = np.sqrt(p["motor_pos"]["xyz"][0] ** 2 * 2) * 2                 # diagonal motor-to-motor distance
        p["propellers"] = {"h": 0.002, "r": self.arm_length / 4, "m": 0.0}
        mxy = self.arm_length * np.sqrt(2) / 4
        p["motor_pos"]["xyz"] = [mxy, mxy, 0.0]
        self.motor_xyz = np.array(p["motor_pos"]["xyz"], dtype=np.float64)
        if "mass" not in p:
            def box_m(d):
                return d["m"] if d.get("m") is not None else d["density"] * d["l"] * d["w"] * d["h"]

            def cyl_m(d):
                return d["m"] if d.get("m") is not None else d["density"] * np.pi * d["h"] * d["r"] ** 2

            p["mass"] = box_m(p["body"]) + box_m(p["payload"]) + 4 * box_m(p["arms"]) + 4 * cyl_m(p["motors"]) + \
                4 * cyl_m(p["propellers"])
        p["arms"] = {"l": self.arm_length, "r": self.arm_length / 20, "m": p["mass"] / 2}
        angle = float(p["arms_pos"]["angle"]) / 180. * np.pi
        self.arm_angle = angle if angle != 0. else 0.01
        arm_xyz = np.array([0., 0., p["arms_pos"]["z"]])
        motors_xyz = (MOTOR_SIGNS * self.motor_xyz[:, None]).T
        rod_m = p["arms"]["m"]
        rod_I = np.diag([rod_m * self.arm_length ** 2 / 12., 0., rod_m * self.arm_length ** 2 / 12.])   # RodLink.I_com :121-126
        self.m = 2 * rod_m                                  # the propellers are massless
        self.com = (2 * rod_m * arm_xyz) / self.m
        total = np.zeros((3, 3))
        for a in (-self.arm_angle, self.arm_angle):
            c, s = np.cos(a), np.sin(a)
            Rz = np.array([[c, -s, 0.], [s, c, 0.], [0., 0., 1.]])
            total = total + _shift(Rz @ rod_I @ Rz.T, rod_m, arm_xyz - self.com)
        self.I_com = total
        self.prop_pos = motors_xyz - self.com

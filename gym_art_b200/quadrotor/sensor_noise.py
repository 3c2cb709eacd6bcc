"""SensorNoise: the reference's sensing-noise *configuration* object (gym_art/quadrotor/sensor_noise.py:57-98).

On the B200 path the noise itself is applied inside the fused env-step kernel (csrc/qs_device.cuh build_obs, which
follows add_noise :100-159, add_noise_to_omega :162-170, quat_from_small_angle :9-21, rot2quat :26-55); this class
carries the constructor arguments into QsSensorNoise.  add_noise() is offered for API compatibility and runs the
same kernel on a one-env handle, it does not re-implement the arithmetic on the CPU.
"""


class SensorNoise(object):
    def __init__(self, pos_norm_std=0.005, pos_unif_range=0., vel_norm_std=0.01, vel_unif_range=0., quat_norm_std=0.,
                 quat_unif_range=0., gyro_norm_std=0., gyro_noise_density=0.000175, gyro_random_walk=0.0105,
                 gyro_bias_correlation_time=1000., bypass=False, acc_static_noise_std=0.002,
                 acc_dynamic_noise_ratio=0.005):
        self.pos_norm_std = pos_norm_std
        self.pos_unif_range = pos_unif_range
        self.vel_norm_std = vel_norm_std
        self.vel_unif_range = vel_unif_range
        self.quat_norm_std = quat_norm_std
        self.quat_unif_range = quat_unif_range
        self.gyro_noise_density = gyro_noise_density
        self.gyro_random_walk = gyro_random_walk
        self.gyro_bias_correlation_time = gyro_bias_correlation_time
        self.gyro_norm_std = gyro_norm_std
        self.acc_static_noise_std = acc_static_noise_std
        self.acc_dynamic_noise_ratio = acc_dynamic_noise_ratio
        self.bypass = bypass

    def add_noise(self, pos, vel, rot, omega, acc, dt, device="cuda:0", seed=None):
        """sensor_noise.py:100-159 for one drone: returns noisy (pos, vel, rot, omega, acc).  Runs the observation stage
        of the CUDA path on a one-env handle (pinned host buffers: one launch + one sync); `rot` must be a 3x3
        matrix (the reference also accepts Euler angles / quaternions, which no observation builder on the step path
        uses)."""
        import numpy as np
        rot = np.asarray(rot, dtype=np.float64)
        if rot.shape != (3, 3):
            raise NotImplementedError("add_noise on the device takes the rotation as a 3x3 matrix")
        if self.bypass:
            return pos, vel, rot, omega, acc
        key = round(1.0 / dt, 9)
        if getattr(self, "_dev", None) is None or self._dev[0] != key:
            from ..vec_env import QuadVecEnv
            if seed is None:
                seed = int(np.random.randint(0, 2 ** 31 - 1))
            kw = {k: v for k, v in self.kwargs().items() if k != "bypass"}
            self._dev = (key, QuadVecEnv(1, sim_freq=1.0 / dt, obs_repr="xyz_vxyz_R_omega_acc_act", sense_noise=kw,
                                         goal=(0., 0., 0.), dtype="float64", device=device, seed=seed, auto_reset=False,
                                         host_mapped=True))
        env = self._dev[1]
        for name, v in (("pos", pos), ("vel", vel), ("omega", omega)):
            env.set_field(name, np.asarray(v, dtype=np.float64).reshape(1, 3))
        env.set_field("rot", rot.reshape(1, 9))
        o = env.observe(np.asarray(acc, dtype=np.float64).reshape(1, 3)).numpy()[0].astype(np.float64)
        return o[0:3], o[3:6], o[6:15].reshape(3, 3), o[15:18], o[18:21]

    def kwargs(self):
        return {k: getattr(self, k) for k in (
            "pos_norm_std", "pos_unif_range", "vel_norm_std", "vel_unif_range", "quat_norm_std", "quat_unif_range",
            "gyro_norm_std", "gyro_noise_density", "gyro_random_walk", "gyro_bias_correlation_time", "bypass",
            "acc_static_noise_std", "acc_dynamic_noise_ratio")}

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

    def kwargs(self):
        return {k: getattr(self, k) for k in (
            "pos_norm_std", "pos_unif_range", "vel_norm_std", "vel_unif_range", "quat_norm_std", "quat_unif_range",
            "gyro_norm_std", "gyro_noise_density", "gyro_random_walk", "gyro_bias_correlation_time", "bypass",
            "acc_static_noise_std", "acc_dynamic_noise_ratio")}

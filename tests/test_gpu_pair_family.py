"""GPU: the one two-group configuration with a kernel of its own - rotor drag / linearity / damping together with every
sensor-noise term (F_DRAG | F_NOISEX, `tma-generic-drag-noise`; everything else with more than one feature group takes the
catch-all) - against the NumPy oracle, fp32, per step from the device's state, injected draws."""
import numpy as np
import pytest

from oracle import quadsim_oracle as qo

pytestmark = pytest.mark.gpu

DRAG = {"motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424}, "damp": {"vel": 0.001, "omega_quadratic": 0.015}}
NOISE = dict(pos_unif_range=0.01, vel_unif_range=0.02, quat_norm_std=0.01, quat_unif_range=0.02, gyro_norm_std=0.1)


@pytest.mark.parametrize("controller,octrl", [("raw", qo.CTRL_RAW), ("mellinger", qo.CTRL_MELLINGER)])
def test_drag_and_full_sensor_noise_family_matches_oracle_fp32(controller, octrl):
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    params = quad_models.crazyflie_params()
    qr._update_existing(params, DRAG)
    n, T = 1500, 30                                              # > 1,184 envs would not matter here: small launches are one wave
    kw = dict(ep_time=0.1, sim_steps=2, sense_noise=NOISE, init_random_state=True)     # 10-step episodes: auto-resets inside the run
    env = QuadVecEnv(n, dynamics_params=params, controller=controller, noise_mode="injected", auto_reset=True, **kw)
    oe = qo.OracleVecEnv(params, n, controller=octrl, auto_reset=True, **kw)
    rng = np.random.RandomState(17)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)

    def obs_draws(width, off):
        z = rng.standard_normal((n, width))
        for s in (qo.Z_UPOS, qo.Z_UVEL, qo.Z_UTHETA):
            z[:, off + s:off + s + 3] = rng.uniform(-1, 1, (n, 3))
        return z

    def reset_draws():
        rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
        rz[:, qo.RZ_NUP:qo.RZ_OBS] = rng.standard_normal((n, qo.RZ_OBS - qo.RZ_NUP))
        rz[:, qo.RZ_OBS:] = obs_draws(qo.OBS_Z, 0)
        return f32(rz)

    rz = reset_draws()
    env.set_noise(reset_z=rz)
    env.reset()
    oe.reset(reset_z=rz)
    fields = (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
              ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou"), ("gyro_bias", "gyro_bias"))
    worst = 0.0
    for t in range(T):
        for f, of in fields:
            setattr(oe, of, env.get_field(f).double().cpu().numpy())
        a = f32(rng.uniform(-0.5, 0.8, (n, 4)))
        ou_z, obs_z, rz = f32(rng.standard_normal((n, 2, 4))), f32(obs_draws(qo.OBS_Z, 0)), reset_draws()
        env.set_noise(ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        o, r, d = env.step(None if controller == "mellinger" else a)
        oo, ro, do = oe.step(np.zeros((n, 4)) if controller == "mellinger" else a, ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        assert np.array_equal(d.cpu().numpy().astype(bool), np.asarray(do).astype(bool))
        err = (np.abs(o.double().cpu().numpy() - oo) / np.maximum(np.abs(oo), 1.0)).max()
        worst = max(worst, err)
        assert err < 1e-5, (t, err)
        np.testing.assert_allclose(r.double().cpu().numpy(), ro, rtol=0, atol=2e-6)
    assert env.kernel_variant == "tma-generic-drag-noise-injected", env.kernel_variant
    print("worst relative observation error over %d steps: %.2e" % (T, worst))
    env.close()

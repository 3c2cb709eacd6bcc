"""GPU: edge cases against the oracle - a single env, sizes around the warp / CTA boundaries, one and many substeps,
two-step episodes (reset on every other step), all-masked and none-masked resets."""
import numpy as np
import pytest

from oracle import quadsim_oracle as qo

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n", [1, 31, 33, 127, 129])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 3e-5)])
def test_ragged_sizes_match_oracle(n, dtype, tol):
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    params = quad_models.crazyflie_params()
    rng = np.random.RandomState(n)
    for ctrl, oc, sim_steps, ep_time in (("raw", qo.CTRL_RAW, 2, 0.01), ("mellinger", qo.CTRL_MELLINGER, 1, 0.02),
                                         ("omega_thrust", qo.CTRL_OMEGA_THRUST, 5, 0.05)):
        kw = dict(sim_steps=sim_steps, ep_time=ep_time, sense_noise="default", init_random_state=True)
        oe = qo.OracleVecEnv(params, n, controller=oc, auto_reset=True, **kw)
        env = QuadVecEnv(n, dynamics_params=params, controller=ctrl, dtype=dtype, noise_mode="injected", auto_reset=True, **kw)
        assert env.ep_len == oe.ep_len and env.ep_len <= 4

        def draws():
            rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
            rz[:, qo.RZ_NUP:] = rng.standard_normal((n, qo.RESET_Z - qo.RZ_NUP))
            return rz

        rz = draws()
        env.set_noise(reset_z=rz)
        o = env.reset().double().cpu().numpy()
        np.testing.assert_allclose(o, oe.reset(reset_z=rz), rtol=0, atol=tol)
        for t in range(9):
            if dtype == "float32":                   # per-step comparison from the device's own (rounded) state
                for f, of in (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
                              ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou")):
                    setattr(oe, of, env.get_field(f).double().cpu().numpy())
            a = rng.uniform(-0.3, 0.6, (n, 4))
            if dtype == "float32":
                a = a.astype(np.float32).astype(np.float64)
            ou_z, obs_z, rz = rng.standard_normal((n, sim_steps, 4)), rng.standard_normal((n, qo.OBS_Z)), draws()
            if dtype == "float32":
                ou_z, obs_z, rz = (x.astype(np.float32).astype(np.float64) for x in (ou_z, obs_z, rz))
            env.set_noise(ou_z=ou_z, obs_z=obs_z, reset_z=rz)
            o, r, d = env.step(a)
            oo, ro, do = oe.step(a, ou_z=ou_z, obs_z=obs_z, reset_z=rz)
            assert np.array_equal(d.cpu().numpy(), do), (ctrl, t)
            err = np.max(np.abs(o.double().cpu().numpy() - oo) / np.maximum(np.abs(oo), 1.0))
            assert err < tol, (ctrl, t, err)
            assert np.max(np.abs(r.double().cpu().numpy() - ro)) < tol
        assert env.stats()[2] == oe.stats[2] > 0
        env.close()


def test_masked_reset_extremes():
    import torch
    from gym_art_b200 import QuadVecEnv
    n = 300
    env = QuadVecEnv(n, seed=3)
    env.reset()
    a = torch.zeros((n, 4), device="cuda")
    for _ in range(3):
        env.step(a)
    before, obs_before = env.state.clone(), env.obs.clone()
    env.reset(torch.zeros(n, dtype=torch.bool, device="cuda"))               # nobody
    assert torch.equal(env.state, before) and torch.equal(env.obs, obs_before)
    env.reset(torch.ones(n, dtype=torch.bool, device="cuda"))                # everybody == reset()
    assert bool((env.tick == 0).all()) and not torch.equal(env.state, before)
    env.close()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 3e-5)])
def test_vertplane_2d_env_matches_oracle(dtype, tol):
    """dim_mode='2D' (VertPlaneControl, two-wide actions, spawn in the x-z plane): the reference env cannot run this
    mode end to end, the oracle pins the controller on the live reference; here the CUDA env
    step against the oracle's env step."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    n = 333
    params = quad_models.defaultquad_params()
    rng = np.random.RandomState(8)
    kw = dict(dim_mode="2D", ep_time=0.15, sense_noise="default", obs_repr="xyz_vxyz_R_omega_h")
    oe = qo.OracleVecEnv(params, n, controller=qo.CTRL_VERTPLANE, auto_reset=True, **kw)
    env = QuadVecEnv(n, dynamics_params=params, dtype=dtype, noise_mode="injected", auto_reset=True, **kw)
    assert env.controller_name == "vertplane" and env.act_dim == 2
    rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
    env.set_noise(reset_z=rz)
    np.testing.assert_allclose(env.reset().double().cpu().numpy(), oe.reset(reset_z=rz), rtol=0, atol=tol)
    assert float(env.pos[:, 1].abs().max()) == 0.0                   # y = goal.y in 2D (quadrotor.py:1091-1092)
    for t in range(40):
        if dtype == "float32":
            for f, of in (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
                          ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou")):
                setattr(oe, of, env.get_field(f).double().cpu().numpy())
        cast = (lambda x: x.astype(np.float32).astype(np.float64)) if dtype == "float32" else (lambda x: x)
        a = cast(rng.uniform(-1, 1, (n, 2)))
        ou_z, obs_z, rz = cast(rng.standard_normal((n, 2, 4))), cast(rng.standard_normal((n, qo.OBS_Z))), cast(rng.uniform(-1, 1, (n, qo.RESET_Z)))
        env.set_noise(ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        assert np.array_equal(d.cpu().numpy(), do)
        err = np.max(np.abs(o.double().cpu().numpy() - oo) / np.maximum(np.abs(oo), 1.0))
        assert err < tol and np.max(np.abs(r.double().cpu().numpy() - ro)) < tol, (t, err)
    env.close()

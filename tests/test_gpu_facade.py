"""GPU: the reference-shaped facades (QuadrotorEnv, QuadrotorDynamics, controllers, quadrotor_multi.QuadrotorEnv)
driven the way code written against gym_art drives them, checked against the fixtures recorded from the live
reference (tests/golden/, oracle/gen_golden.py)."""
import numpy as np
import pytest

from tests.golden_util import Golden, load_npz

pytestmark = pytest.mark.gpu


def test_known_answer_through_the_reference_api():
    """SURVEY.md section 8c known-answer case, written exactly as against the reference objects."""
    from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv
    ka = load_npz("known_answer")
    for tag, quad in (("hb", "DefaultQuad"), ("cf", "Crazyflie")):
        env = QuadrotorEnv(dynamics_params=quad, raw_control=True, ep_time=5,
                           dynamics_change={"noise": {"thrust_noise_ratio": 0.0}}, seed=1)
        env.dynamics.set_state(np.array([0.5, -0.25, 1.5]), np.array([0.1, 0.2, -0.3]), ka["R0"], np.array([0.5, -0.4, 0.3]))
        env.dynamics.reset()
        env.dynamics.since_last_svd = 0
        a = np.array([0.1, -0.2, 0.3, 0.0])
        for t in range(100):
            o, r, d, info = env.step(a)
            if t in (0, 99):
                dyn = env.dynamics
                for k, v in (("pos", dyn.pos), ("vel", dyn.vel), ("rot", dyn.rot), ("omega", dyn.omega),
                             ("cmds_damp", dyn.thrust_cmds_damp), ("obs", o)):
                    np.testing.assert_allclose(v, ka["%s_%d_%s" % (tag, t + 1, k)], rtol=0, atol=1e-9, err_msg="%s %s" % (tag, k))
                assert abs(r - float(ka["%s_%d_rew" % (tag, t + 1)])) < 1e-12
                assert env.crashed == bool(ka["%s_%d_crashed" % (tag, t + 1)])
                assert not d and env.tick == t + 1
                # info dict (quadrotor.py:993-1028): weighted costs sum to the reward, raw pos cost is the distance
                rw = info["rewards"]
                total = sum(rw[k] for k in rw if k.startswith("rew_") and k != "rew_main")
                assert abs(env.dt * total - r) < 1e-12
                assert abs(-rw["rewraw_pos"] - np.linalg.norm(env.goal - dyn.pos)) < 1e-12
                np.testing.assert_allclose(info["obs_comp"]["act_filtered"][0], dyn.thrust_cmds_damp)
                assert info["dyn_params"]["dt"][0] == env.dt * env.sim_steps
        assert tuple(env.observation_space.shape) == (18,) and tuple(env.action_space.shape) == (4,)
        env.close()


@pytest.mark.parametrize("name,module", [("hb_mellinger_noise_act", "quadrotor"), ("cf_raw_noise_rnd", "quadrotor"),
                                         ("hb_multi_reward", "multi"), ("hb_mellinger_excite", "quadrotor"),
                                         ("cf_raw_resample_goal", "quadrotor")])
def test_env_facade_replays_reference_rollout(name, module):
    """reset()/step() through the facade with the reference's recorded noise tape injected: obs / reward / done of
    every step equal the reference's to 1e-9 over whole episodes (fp64 build)."""
    g = Golden(name)
    if module == "multi":
        from gym_art_b200.quadrotor_multi.quadrotor_multi import QuadrotorEnv
    else:
        from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv
    kw = dict(g.meta["env_kwargs"])
    env = QuadrotorEnv(noise_mode="injected", seed=0, **kw)
    assert env.ep_len == g.meta["ep_len"]
    vec = env._vec
    vec.clock = (float(g.svd_timer0), 0)
    if vec.layout.row_ou >= 0:
        vec.set_field("ou_state", g.ou0)
    vec.set_noise(reset_z=g.reset_z[0][None], goal_z=g.reset_goal_z[0][None])
    o = env.reset()
    np.testing.assert_allclose(o, g.reset_obs[0], rtol=0, atol=1e-9)
    T = min(g.T, 620)
    for t in range(T):
        vec.set_noise(ou_z=g.ou_z[t][None], obs_z=g.obs_z[t][None], goal_z=g.goal_z[t][None])
        o, r, d, info = env.step(g.actions[t])
        if "goal" in g.a:
            np.testing.assert_allclose(env.goal, g.goal[t], rtol=0, atol=1e-12)
        np.testing.assert_allclose(o, g.obs[t], rtol=0, atol=2e-9, err_msg="obs step %d" % t)
        assert abs(r - g.rew[t]) < 1e-10 and d == bool(g.done[t])
        np.testing.assert_allclose(env.dynamics.accelerometer, g.accelerometer[t], rtol=0, atol=1e-8)
        if d:
            k = g.reset_index_after(t)
            vec.set_noise(reset_z=g.reset_z[k][None], goal_z=g.reset_goal_z[k][None])
            np.testing.assert_allclose(env.reset(), g.reset_obs[k], rtol=0, atol=1e-9)
            assert env.tick == 0 and env.traj_count == k
    env.close()


@pytest.mark.parametrize("name,cls", [("hb_velyaw", "VelocityYawControl"), ("cf_omegathrust_noise_rnd", "OmegaThrustControl"),
                                      ("hb_shifted", "ShiftedMotorControl")])
def test_controllers_on_standalone_dynamics(name, cls):
    """controller.step(dynamics, action, dt) on a QuadrotorDynamics built by hand, as the reference's users do
    (these three controllers are never selected by QuadrotorEnv itself)."""
    from gym_art_b200.quadrotor import quadrotor_control as qc
    from gym_art_b200.quadrotor.quadrotor import QuadrotorDynamics
    g = Golden(name)
    ok = g.oracle_kwargs
    dyn = QuadrotorDynamics(g.params, dynamics_steps_num=ok["sim_steps"], noise_mode="injected")
    ctrl = getattr(qc, cls)(dyn)
    dt = 1.0 / ok["sim_freq"]
    assert ctrl.action_space(dyn).shape == (4,)
    checked = 0
    for t in range(0, min(g.T - 1, 120)):
        if g.done[t] or g.done[t + 1]:
            continue
        dyn.set_state(g.pos[t], g.vel[t], g.rot[t], g.omega[t])
        dyn.omega = g.omega[t]                                  # mid-episode omega is float64 (only reset casts, :223)
        dyn._vec.set_field("thrust_rot_damp", g.rot_damp[t])
        dyn._vec.set_field("thrust_cmds_damp", g.cmds_damp[t])
        if dyn._vec.layout.row_ou >= 0:
            dyn._vec.set_field("ou_state", g.ou[t])
        dyn.since_last_svd = 0.0
        dyn._vec.set_noise(ou_z=g.ou_z[t + 1][None])
        ctrl.step(dyn, g.actions[t + 1], dt)
        fired = (t + 2) * ok["sim_steps"] * dt % 0.5 < ok["sim_steps"] * dt + 1e-9   # reference re-orthonormalised here
        tol = 1e-6 if fired else 1e-10
        for k, v in (("pos", dyn.pos), ("vel", dyn.vel), ("rot", dyn.rot), ("omega", dyn.omega), ("cmds_damp", dyn.thrust_cmds_damp)):
            np.testing.assert_allclose(v, g.a[k][t + 1], rtol=0, atol=tol, err_msg="%s step %d" % (k, t))
        assert np.all(ctrl.action >= 0) and np.all(ctrl.action <= 1)
        checked += 1
    assert checked > 50
    dyn.close()


def test_dynamics_object_api():
    from gym_art_b200.quadrotor.quadrotor import QuadrotorDynamics
    from gym_art_b200.quadrotor import quad_models
    dyn = QuadrotorDynamics(quad_models.crazyflie_params(), dynamics_steps_num=2)
    assert abs(dyn.mass - 0.028) < 1e-9 and dyn.inertia.shape == (3,) and dyn.prop_pos.shape == (4, 3)
    pos, vel, rot, omega = dyn.random_state(box=3.0, vel_max=2.0, omega_max=1.0)
    assert np.all(np.abs(pos) <= 3.0) and np.linalg.norm(vel) <= 2.0 + 1e-9 and np.linalg.norm(omega) <= 1.0 + 1e-9
    np.testing.assert_allclose(rot @ rot.T, np.eye(3), atol=1e-12)
    dyn.set_state(pos, vel, rot, omega)
    np.testing.assert_array_equal(dyn.omega, omega.astype(np.float32).astype(np.float64))   # quadrotor.py:223
    np.testing.assert_array_equal(dyn.accelerometer, [0, 0, 9.81])
    dyn.step(np.array([0.5, 0.5, 0.5, 0.5]), 0.005)
    assert dyn.state_vector().shape == (18,) and np.all(dyn.thrust_cmds_damp > 0)
    hover = dyn.rot.T @ (dyn.acc + np.array([0, 0, 9.81]))
    np.testing.assert_allclose(dyn.accelerometer, hover, atol=1e-12)
    dyn.step(np.array([0.5, 0.5, 0.5, 0.5]), 0.01)              # another dt: a second handle over the same drone
    assert np.all(np.isfinite(dyn.pos))
    dyn.close()


def test_nan_reward_raises_value_error_like_the_reference():
    from gym_art_b200.quadrotor.quadrotor import QuadrotorEnv
    env = QuadrotorEnv(seed=2)
    env.dynamics.pos = np.array([np.nan, 0.0, 1.0])
    with pytest.raises(ValueError, match="reward is Nan"):
        env.step(np.zeros(4))
    env.close()

"""GPU: the CUDA path (through the C-ABI, via QuadVecEnv) against the fixtures recorded from the live reference
and against the NumPy oracle on the same seeded inputs.

Bars (BASELINE.json north_star):
  * fp64 build: 1e-12 per env step (state re-synchronised to the reference before every step), and a
    whole-rollout bound that only allows the round-off growth the float64 reference itself shows.
  * fp32 build: 1e-5 relative per step (|err| <= 1e-5 * max(|x|, floor)), and a drift bound over 500-step
    rollouts stated separately for closed-loop (Mellinger) and open-loop (random actions) runs.
"""
import numpy as np
import pytest

from oracle import quadsim_oracle as qo
from tests.golden_util import Golden, ROLLOUT_CASES, load_npz, obs_tolerance

pytestmark = pytest.mark.gpu

CTRL_NAMES = {qo.CTRL_RAW: "raw", qo.CTRL_VERTICAL: "vertical", qo.CTRL_VERTPLANE: "vertplane", qo.CTRL_SHIFTED: "shifted",
              qo.CTRL_OMEGA_THRUST: "omega_thrust", qo.CTRL_VELOCITY_YAW: "velocity_yaw", qo.CTRL_MELLINGER: "mellinger"}

STATE_FIELDS = (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
                ("thrust_cmds_damp", "cmds_damp"))


# which kernel the fp32 build selects for each fixture: the specialised TMA kernels wherever the configuration is in
# the fast subset, the generic kernel for rare features (drag / linearity / damping, full sensor noise, 1D mode, ...)
EXPECTED_VARIANT = {
    "hb_raw_nonoise": "tma-raw-injected", "cf_raw_noise_rnd": "tma-raw-injected", "hb_raw_nozm": "tma-raw-injected",
    "hb_raw_allrew": "tma-raw-injected", "hb_multi_reward": "tma-raw-injected", "hb_xyzr_h_noise": "tma-raw-injected",
    "mq_raw_simsteps4": "tma-raw-injected", "rq_raw_noise": "tma-raw-injected",
    "cf_omegathrust_noise_rnd": "tma-omega-thrust-injected", "cf_relative_omegathrust": "tma-omega-thrust-injected",
    "hb_mellinger_noise_act": "tma-mellinger-injected", "hb_mellinger_nonoise": "tma-mellinger-injected",
    "cf_mellinger_rnd": "tma-mellinger-injected", "rq_mellinger": "tma-mellinger-injected",
    "hb_velyaw": "tma-velocity-yaw-injected", "hb_shifted": "tma-shifted-injected", "cf_raw_simplified": "tma-raw-injected",
    # one rarely used feature group at sim_steps = 2: the kernel that carries just that group (qs::Feat)
    "cf_drag_damp": "tma-generic-drag-injected", "cf_fullnoise_accact": "tma-raw-injected",       # F_NOISEX with RawControl: compile-time-controller twin
    "hb_vertical": "tma-generic-misc-injected", "hb_vertical_rnd": "tma-generic-misc-injected",
    # ... F_MISC with the RawControl / Mellinger controllers: the compile-time-controller twins of that group
    "hb_mellinger_excite": "tma-mellinger-injected", "cf_raw_resample_goal": "tma-raw-injected",
    "cf_xyzr_quat_h_rnd": "tma-raw-injected", "hb_t2w_t2t_nonoise": "tma-raw-injected",
    "hb_xyzr_t2w_mellinger": "tma-mellinger-injected",
    "hb_quat_fullnoise": "tma-generic-injected",          # two groups (quaternion builder + full noise): the catch-all
}


def make_env(g, n, dtype, noise_mode="injected", auto_reset=False, info=False, **over):
    from gym_art_b200 import QuadVecEnv
    ok = g.oracle_kwargs
    kw = dict(dynamics_params=g.params, controller=CTRL_NAMES[g.controller], raw_control_zero_middle=ok["zero_middle"],
              dim_mode=ok["dim_mode"], sim_freq=ok["sim_freq"], sim_steps=ok["sim_steps"], obs_repr=ok["obs_repr"],
              ep_time=ok["ep_time"], room_size=ok["room_size"], init_random_state=ok["init_random_state"],
              rew_coeff=ok["rew_coeff"], sense_noise=ok["sense_noise"], gravity=ok["gravity"],
              multi_reward=ok["multi_reward"], dtype=dtype, noise_mode=noise_mode, auto_reset=auto_reset, info=info,
              resample_goal=ok.get("resample_goal", False), excite=ok.get("excite", False),
              dynamics_simplification=ok.get("dynamics_simplification", False), t2w_std=ok.get("t2w_std", 0.005),
              t2t_std=ok.get("t2t_std", 0.0005))
    kw.update(over)
    env = QuadVecEnv(n, **kw)
    env.clock = (float(g.svd_timer0), 0)
    if env.layout.row_ou >= 0:
        env.set_field("ou_state", g.ou0)
    if env.layout.row_gyro_bias >= 0:
        env.set_field("gyro_bias", g.gyro_bias0)
    return env


def tile(x, n):
    return np.tile(np.asarray(x)[None], (n,) + (1,) * np.ndim(x))


def state_of(env):
    return {gk: getattr(env, ek).double().cpu().numpy() for ek, gk in STATE_FIELDS}


def resync(env, g, t):
    """Put the reference's state after step t into every env (t = -1: nothing to do, the reset did it)."""
    env.set_field("pos", g.pos[t])
    env.set_field("vel", g.vel[t])
    env.set_field("rot", g.rot[t].reshape(9))
    env.set_field("omega", g.omega[t])
    env.set_field("thrust_rot_damp", g.rot_damp[t])
    env.set_field("thrust_cmds_damp", g.cmds_damp[t])
    if env.layout.row_ou >= 0:
        env.set_field("ou_state", g.ou[t])
    if env.layout.row_goal >= 0 and "goal" in g.a:
        env.set_field("goals", g.goal[t])


def rel_err(a, b, floor):
    return np.max(np.abs(a - b) / np.maximum(np.abs(b), floor))


def obs_err(g, o, ref, relative=True):
    """Observation error against the reference row, relative to max(|x|, 1) (or absolute); the quaternion builders' four
    components are weighed by the conditioning of the reference's single-branch R2quat (tests/golden_util.obs_tolerance)."""
    w = obs_tolerance(g.oracle_kwargs["obs_repr"], np.asarray(ref), 1.0)
    return np.max(np.abs(o - ref) / (np.maximum(np.abs(ref), 1.0) if relative else 1.0) / w)


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_fp64_rollout_matches_reference(name):
    """Free-running fp64 rollout vs the recorded reference rollout: obs, reward, done, state, resets."""
    g = Golden(name)
    n = 3
    env = make_env(g, n, "float64")
    A = env.act_dim
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    obs = env.reset().cpu().numpy()
    assert obs_err(g, obs, tile(g.reset_obs[0], n), False) <= 1e-12
    worst = 0.0
    for t in range(g.T):
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), goal_z=tile(g.goal_z[t], n))
        o, r, d = env.step(tile(g.actions[t][:A], n))
        o, r, d = o.cpu().numpy(), r.cpu().numpy(), d.cpu().numpy()
        assert bool(d[0]) == bool(g.done[t]) and np.all(d == d[0])
        st = state_of(env)
        err = max(obs_err(g, o, g.obs[t], False), np.abs(r - g.rew[t]).max(),
                  max(np.abs(st[k] - g.a[k][t]).max() for k in st))
        worst = max(worst, err)
        assert err < 2e-10, "step %d err %g" % (t, err)
        if d[0]:
            i = g.reset_index_after(t)
            env.set_noise(reset_z=tile(g.reset_z[i], n), goal_z=tile(g.reset_goal_z[i], n))
            ob = env.reset().cpu().numpy()
            assert obs_err(g, ob, tile(g.reset_obs[i], n), False) <= 1e-12
    env.check_finite()
    print(name, "fp64 free-running worst abs err", worst)


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_fp64_per_step_1e12(name):
    """north_star: 'an fp64 build must match to 1e-12 per step' - state re-synchronised before every step."""
    g = Golden(name)
    n = 2
    env = make_env(g, n, "float64")
    A = env.act_dim
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    env.reset()
    for t in range(min(g.T, 400)):
        if t > 0 and not g.done[t - 1]:
            resync(env, g, t - 1)
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), goal_z=tile(g.goal_z[t], n))
        o, r, d = env.step(tile(g.actions[t][:A], n))
        st = state_of(env)
        scale = lambda x: np.maximum(np.abs(x), 1.0)
        assert obs_err(g, o.cpu().numpy(), g.obs[t]) < 1e-12, "obs step %d" % t
        assert np.max(np.abs(r.cpu().numpy() - g.rew[t])) < 1e-12, "rew step %d" % t
        for k in st:
            assert np.max(np.abs(st[k] - g.a[k][t]) / scale(g.a[k][t])) < 1e-12, "%s step %d" % (k, t)
        if g.done[t]:
            i = g.reset_index_after(t)
            env.set_noise(reset_z=tile(g.reset_z[i], n), goal_z=tile(g.reset_goal_z[i], n))
            ob = env.reset().cpu().numpy()
            assert obs_err(g, ob, tile(g.reset_obs[i], n), False) <= 1e-12


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_fp32_per_step_1e5_relative(name):
    """north_star: 'the default fp32 build must stay within 1e-5 relative error per substep'.
    Checked per env step (sim_steps substeps) from the reference's state: |err| <= 1e-5 * max(|x|, floor) with
    floor 1 for pos/vel/omega/obs, 1 for R entries and motor states (bounded by 1), 0.01 for the reward."""
    g = Golden(name)
    n = 2
    env = make_env(g, n, "float32")
    A = env.act_dim
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    ob = env.reset().cpu().numpy()
    assert obs_err(g, ob, g.reset_obs[0]) < 1e-5
    worst = 0.0
    for t in range(min(g.T, 400)):
        if t > 0 and not g.done[t - 1]:
            resync(env, g, t - 1)
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), goal_z=tile(g.goal_z[t], n))
        o, r, d = env.step(tile(g.actions[t][:A], n))
        assert bool(d[0].item()) == bool(g.done[t])
        st = state_of(env)
        e = max(obs_err(g, o.double().cpu().numpy(), g.obs[t]), rel_err(r.double().cpu().numpy(), g.rew[t], 0.01),
                max(rel_err(st[k], g.a[k][t], 1.0) for k in st))
        worst = max(worst, e)
        assert e < 1e-5, "step %d rel err %g" % (t, e)
        if g.done[t]:
            i = g.reset_index_after(t)
            env.set_noise(reset_z=tile(g.reset_z[i], n), goal_z=tile(g.reset_goal_z[i], n))
            ob = env.reset().cpu().numpy()
            assert obs_err(g, ob, g.reset_obs[i]) < 1e-5
    print(name, "fp32 per-step worst rel err", worst)
    # the throughput kernels themselves (injected-noise twins) are what this test holds to the bar
    assert env.kernel_variant == EXPECTED_VARIANT[name], (name, env.kernel_variant)


@pytest.mark.parametrize("name,bound", [("hb_mellinger_nonoise", 1e-4), ("hb_mellinger_noise_act", 1e-4),
                                        ("hb_raw_nonoise", 5e-3), ("cf_omegathrust_noise_rnd", 5e-3)])
def test_fp32_500_step_drift(name, bound):
    """Stated drift bound of the free-running fp32 build over one 500-step episode, relative to magnitude
    (max(|x|, 1)): closed loop (Mellinger) 1e-4; open loop with random actions (tumbling, crashed, |vel| of tens
    of m/s) 5e-3.  The open-loop figure is chaotic amplification of fp32 round-off, not a formula error: the
    per-step test above bounds that."""
    g = Golden(name)
    n = 2
    env = make_env(g, n, "float32")
    A = env.act_dim
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    env.reset()
    worst = 0.0
    for t in range(min(g.T, 500)):
        if g.done[t]:
            break
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), goal_z=tile(g.goal_z[t], n))
        env.step(tile(g.actions[t][:A], n))
        st = state_of(env)
        worst = max(worst, max(rel_err(st[k], g.a[k][t], 1.0) for k in ("pos", "vel", "rot", "omega")))
    print(name, "fp32 500-step drift", worst)
    assert worst < bound


def test_known_answer_fp64_and_fp32():
    """SURVEY.md section 8c known-answer state (regenerated from the live reference by oracle/gen_golden.py)."""
    import json
    from gym_art_b200 import QuadVecEnv
    ka = load_npz("known_answer")
    for tag in ("hb", "cf"):
        params = json.loads(str(ka["%s_params_json" % tag]))
        for dtype, tol in (("float64", 1e-11), ("float32", 2e-4)):
            env = QuadVecEnv(4, dynamics_params=params, ep_time=5, dtype=dtype, auto_reset=False)
            env.set_state([0.5, -0.25, 1.5], [0.1, 0.2, -0.3], ka["R0"], [0.5, -0.4, 0.3])
            a = np.tile(np.array([[0.1, -0.2, 0.3, 0.0]]), (4, 1))
            for t in range(100):
                o, r, d = env.step(a)
                if t + 1 in (1, 100):
                    for ek, gk in (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"),
                                   ("thrust_cmds_damp", "cmds_damp")):
                        got = getattr(env, ek).double().cpu().numpy()[0]
                        want = ka["%s_%d_%s" % (tag, t + 1, gk)]
                        assert rel_err(got, want, 1.0) < tol, (tag, dtype, t + 1, ek, got, want)
                    assert abs(r[0].item() - float(ka["%s_%d_rew" % (tag, t + 1)])) < tol


def test_auto_reset_equals_explicit_reset_fp64():
    g = Golden("mq_raw_simsteps4")
    n = 5
    env = make_env(g, n, "float64", auto_reset=True)
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    env.reset()
    for t in range(g.T):
        rz = None
        if g.done[t]:
            rz = tile(g.reset_z[g.reset_index_after(t)], n)
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), reset_z=rz)
        o, r, d = env.step(tile(g.actions[t], n))
        want = g.reset_obs[g.reset_index_after(t)] if g.done[t] else g.obs[t]
        np.testing.assert_allclose(o.cpu().numpy(), tile(want, n), rtol=0, atol=2e-10)
        np.testing.assert_allclose(r.cpu().numpy(), g.rew[t], rtol=0, atol=2e-10)
        assert bool(d[0].item()) == bool(g.done[t])
    st = env.stats()
    assert st[2] == 5 * n


@pytest.mark.parametrize("name,dtype,tol", [("mq_raw_simsteps4", "float64", 2e-10), ("cf_raw_noise_rnd", "float64", 2e-9),
                                            ("hb_raw_nonoise", "float32", 5e-3)])
def test_terminal_observation_is_the_reference_done_step_obs(name, dtype, tol):
    """With auto-reset the step returns the next episode's first observation; the terminal_obs buffer keeps what the
    reference's step() returns on its done step (quadrotor.py:988) for exactly the envs that finished."""
    g = Golden(name)
    n = 3
    env = make_env(g, n, dtype, auto_reset=True, terminal_obs=True)
    env.set_noise(reset_z=tile(g.reset_z[0], n), goal_z=tile(g.reset_goal_z[0], n))
    env.reset()
    sentinel = 12345.0
    seen = 0
    for t in range(g.T):
        rz = gz = None
        if g.done[t]:
            k = g.reset_index_after(t)
            rz, gz = tile(g.reset_z[k], n), tile(g.reset_goal_z[k], n)
        env.terminal_obs.fill_(sentinel)
        env.set_noise(ou_z=tile(g.ou_z[t], n), obs_z=tile(g.obs_z[t], n), reset_z=rz, goal_z=gz if gz is not None else tile(g.goal_z[t], n))
        o, r, d = env.step(tile(g.actions[t][:env.act_dim], n))
        term = env.terminal_obs.double().cpu().numpy()
        if g.done[t]:
            seen += 1
            assert np.max(np.abs(term - g.obs[t]) / np.maximum(np.abs(g.obs[t]), 1.0)) < tol, t
            want = g.reset_obs[g.reset_index_after(t)]
            assert np.max(np.abs(o.double().cpu().numpy() - want) / np.maximum(np.abs(want), 1.0)) < max(tol, 1e-9)
        else:
            assert np.all(term == sentinel)
    assert seen >= 1
    env.close()


def test_batched_oracle_parity_ragged_n():
    """Many distinct envs (N not a multiple of the CTA size) against the batched oracle, fp64, injected noise,
    auto-reset, sensor noise, random init: every env has its own actions and draws."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    n, T = 1000 + 37, 60
    rng = np.random.RandomState(5)
    params = quad_models.crazyflie_params()
    kw = dict(sim_steps=2, obs_repr="xyz_vxyz_R_omega_acc_act", ep_time=0.2, init_random_state=True, sense_noise="default")
    oe = qo.OracleVecEnv(params, n, controller=qo.CTRL_OMEGA_THRUST, auto_reset=True, **kw)
    env = QuadVecEnv(n, dynamics_params=params, controller="omega_thrust", dtype="float64", noise_mode="injected",
                     auto_reset=True, **kw)
    rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
    rz[:, qo.RZ_NUP:] = rng.standard_normal((n, qo.RESET_Z - qo.RZ_NUP))
    env.set_noise(reset_z=rz)
    np.testing.assert_allclose(env.reset().cpu().numpy(), oe.reset(reset_z=rz), rtol=0, atol=1e-12)
    for t in range(T):
        a = np.concatenate([rng.uniform(-0.2, 0.2, (n, 1)), rng.uniform(-1, 1, (n, 3))], axis=1)
        ou_z = rng.standard_normal((n, 2, 4))
        obs_z = rng.standard_normal((n, qo.OBS_Z))
        obs_z[:, [3, 4, 5, 9, 10, 11, 21, 22, 23]] = rng.uniform(-1, 1, (n, 9))
        rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
        rz[:, qo.RZ_NUP:] = rng.standard_normal((n, qo.RESET_Z - qo.RZ_NUP))
        env.set_noise(ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        assert np.array_equal(d.cpu().numpy(), do)
        np.testing.assert_allclose(o.cpu().numpy(), oo, rtol=0, atol=1e-9, err_msg="obs step %d" % t)
        np.testing.assert_allclose(r.cpu().numpy(), ro, rtol=0, atol=1e-10)
    np.testing.assert_allclose(env.stats(), oe.stats, rtol=1e-9)

"""GPU: SURVEY.md section 8 row f4 - the affine model F(s, dt) / G(s) (quadrotor.py:461-518), the pitch/roll-restricted
spawn law (:243-259), init_random_state in the 1D / 2D modes (:1101-1110) and the quaternion / t2w / t2t observation builders
(get_state.py:276-382) - through the C-ABI against fixtures recorded from the live reference (tests/golden/affine_answers.npz,
restricted_state.npz, the *_quat_*, *_t2w_* and hb_vertical_rnd rollouts run in tests/test_gpu_parity.py) and against the
oracle on seeded inputs.
"""
import json

import numpy as np
import pytest

from oracle import philox as ph
from oracle import quadsim_oracle as qo
from tests.golden_util import load_npz, obs_tolerance

pytestmark = pytest.mark.gpu


def rel_err(a, b, floor=1.0):
    return float(np.max(np.abs(a - b) / np.maximum(np.abs(b), floor)))


def cancel_floor(S, dt, dtype, tol):
    """dV = ((1 - vel_damp) V - V) / dt (quadrotor.py:475) cancels to -vel_damp V / dt: whatever the precision, the result
    carries the rounding of the two O(|V|) terms divided by dt.  fp32: allow 4 ulp of V over dt there; fp64: the plain bar."""
    floor = np.ones((len(S), 21))
    if dtype == "float32":
        floor[:, 3:6] = np.maximum(1.0, np.abs(S[:, 3:6]) * (4 * 2.0 ** -23 / dt) / tol)
    return floor


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-5)])
def test_affine_F_G_match_the_reference(dtype, tol):
    """Both shipped models with damping switched on, all three dim_modes (control_mx 4x1 / 4x2 / identity), 16 random states
    each; relative to max(|x|, 1).  Also: a ragged batch through the same kernel equals the oracle."""
    from gym_art_b200 import QuadVecEnv
    ans = load_npz("affine_answers")
    dt = float(ans["dt"])
    for tag in ("hb", "cf"):
        params = json.loads(str(ans["%s_params_json" % tag]))
        S = ans["%s_s" % tag]
        for dm in ("1D", "2D", "3D"):
            env = QuadVecEnv(4, dynamics_params=params, dim_mode=dm, dtype=dtype, auto_reset=False)
            F, G = env.affine_fg(S, dt)
            F, G = F.double().cpu().numpy(), G.double().cpu().numpy()
            wantF, wantG = ans["%s_%s_F" % (tag, dm)], ans["%s_%s_G" % (tag, dm)]
            assert G.shape == wantG.shape
            floorF = cancel_floor(S, dt, dtype, tol)
            assert np.max(np.abs(F - wantF) / np.maximum(np.abs(wantF), floorF)) < tol, (tag, dm)
            assert rel_err(G, wantG) < tol, (tag, dm)
            before = env.state.clone()
            Fo, _ = env.affine_fg(S[:3], dt, want_G=False)            # F alone / G alone; the bound state is not touched
            _, Go = env.affine_fg(S[:3], dt, want_F=False)
            assert np.array_equal(Fo.double().cpu().numpy(), F[:3]) and np.array_equal(Go.double().cpu().numpy(), G[:3])
            assert bool((env.state == before).all())
            env.close()
    # ragged batch (not a multiple of the CTA size) against the oracle restatement
    from gym_art_b200.quadrotor import quad_models
    params = quad_models.crazyflie_params()
    params["damp"]["omega_quadratic"], params["damp"]["vel"] = 0.02, 0.003
    rng = np.random.RandomState(12)
    S = rng.uniform(-3, 3, (1000 + 37, 21))
    if dtype == "float32":
        S = S.astype(np.float32).astype(np.float64)
    m = qo.model_constants(params)
    env = QuadVecEnv(4, dynamics_params=params, dtype=dtype, auto_reset=False)
    F, G = env.affine_fg(S, 0.005)
    floorF = cancel_floor(S, 0.005, dtype, tol)
    assert np.max(np.abs(F.double().cpu().numpy() - qo.affine_F(m, S, 0.005)) / np.maximum(np.abs(qo.affine_F(m, S, 0.005)), floorF)) < tol
    assert rel_err(G.double().cpu().numpy(), qo.affine_G(m, S, "3D")) < tol
    env.close()


def test_facade_F_G_and_restricted_random_state():
    """The reference-shaped object: dynamics.F(s, dt), dynamics.G(s) on one state vector; pitch_roll_restricted_random_state
    returns (pos, vel, rot, omega) inside its bounds with rot = Rz Ry Rx of bounded Euler angles."""
    from gym_art_b200.quadrotor.quadrotor import QuadrotorDynamics
    ans = load_npz("affine_answers")
    params = json.loads(str(ans["hb_params_json"]))
    for dm in ("1D", "3D"):
        d = QuadrotorDynamics(model_params=params, dynamics_steps_num=2, dim_mode=dm)
        s = ans["hb_s"][5]
        np.testing.assert_allclose(d.F(s, float(ans["dt"])), ans["hb_%s_F" % dm][5], rtol=0, atol=1e-11)
        np.testing.assert_allclose(d.G(s), ans["hb_%s_G" % dm][5], rtol=0, atol=1e-11)
        assert d.G(ans["hb_s"][:4]).shape == (4, 21, d.control_mx.shape[1])
        d.close()
    d = QuadrotorDynamics(model_params=params, dynamics_steps_num=2, num_envs=256)
    pos, vel, rot, omega = d.pitch_roll_restricted_random_state(box=2.0, vel_max=3.0, omega_max=5.0, pitch_max=0.4, roll_max=0.7,
                                                                yaw_max=2.5)
    assert np.all(np.abs(pos) <= 2.0) and np.all(np.linalg.norm(vel, axis=1) <= 3.0 + 1e-9) and np.all(np.linalg.norm(omega, axis=1) <= 5.0 + 1e-6)
    np.testing.assert_allclose(rot @ np.transpose(rot, (0, 2, 1)), np.tile(np.eye(3), (256, 1, 1)), atol=1e-12)
    pitch = -np.arcsin(rot[:, 2, 0])                                     # R = Rz(yaw) Ry(pitch) Rx(roll): R20 = -sin(pitch)
    roll = np.arctan2(rot[:, 2, 1], rot[:, 2, 2])
    yaw = np.arctan2(rot[:, 1, 0], rot[:, 0, 0])
    assert np.all(np.abs(pitch) <= 0.4 + 1e-9) and np.all(np.abs(roll) <= 0.7 + 1e-9) and np.all(np.abs(yaw) <= 2.5 + 1e-9)
    assert np.abs(pitch).max() > 0.3 and np.abs(roll).max() > 0.55 and np.abs(yaw).max() > 2.0   # ... and fills them
    d.close()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-12), ("float32", 1e-5)])
def test_pitch_roll_restricted_reset_matches_the_reference(dtype, tol):
    """QS_INIT_PITCH_ROLL_RESTRICTED fed with the unit draws the live reference consumed (tests/golden/restricted_state.npz):
    velocity, attitude and body rates of the spawned state."""
    from gym_art_b200 import QuadVecEnv
    z = load_npz("restricted_state")
    params = json.loads(str(z["params_json"]))
    n = z["reset_z"].shape[0]
    env = QuadVecEnv(n, dynamics_params=params, dtype=dtype, noise_mode="injected", auto_reset=False,
                     init_random_state="pitch_roll_restricted", max_init_vel=float(z["vel_max"]), max_init_omega=float(z["omega_max"]),
                     pitch_max=float(z["pitch_max"]), roll_max=float(z["roll_max"]), yaw_max=float(z["yaw_max"]))
    env.set_noise(reset_z=z["reset_z"])
    env.reset()
    assert rel_err(env.vel.double().cpu().numpy(), z["vel"]) < tol
    assert rel_err(env.rot.double().cpu().numpy(), z["rot"]) < tol
    assert rel_err(env.omega.double().cpu().numpy(), z["omega"].astype(np.float32).astype(np.float64)) < max(tol, 1e-7)   # :223
    env.close()


@pytest.mark.parametrize("dim_mode,ctrl,adim", [("1D", qo.CTRL_VERTICAL, 1), ("2D", qo.CTRL_VERTPLANE, 2)])
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-10), ("float32", 3e-5)])
def test_random_init_in_1d_and_2d_env_matches_oracle(dim_mode, ctrl, adim, dtype, tol):
    """init_random_state outside 3D (quadrotor.py:1101-1110): the reference's 2D env cannot step, its resets and the 1D
    rollout pin the oracle live; here the CUDA env (auto-reset, so the law is drawn many times) against the oracle."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    n = 257
    params = quad_models.defaultquad_params()
    rng = np.random.RandomState(9)
    kw = dict(dim_mode=dim_mode, ep_time=0.1, sense_noise="default", init_random_state=True)
    oe = qo.OracleVecEnv(params, n, controller=ctrl, auto_reset=True, **kw)
    env = QuadVecEnv(n, dynamics_params=params, dtype=dtype, noise_mode="injected", auto_reset=True, **kw)
    assert env.act_dim == adim
    cast = (lambda x: x.astype(np.float32).astype(np.float64)) if dtype == "float32" else (lambda x: x)
    rz = cast(rng.uniform(-1, 1, (n, qo.RESET_Z)))
    env.set_noise(reset_z=rz)
    np.testing.assert_allclose(env.reset().double().cpu().numpy(), oe.reset(reset_z=rz), rtol=0, atol=tol)
    if dim_mode == "1D":
        assert float(env.vel[:, :2].abs().max()) == 0.0 and float(env.vel[:, 2].min()) >= 0.0
    else:
        assert float(env.vel[:, 1].abs().max()) == 0.0 and float(env.omega[:, 1].min()) >= 0.0
    for t in range(30):
        if dtype == "float32":
            for f, of in (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
                          ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou")):
                setattr(oe, of, env.get_field(f).double().cpu().numpy())
        a = cast(rng.uniform(-1, 1, (n, adim)))
        ou_z, obs_z, rz = cast(rng.standard_normal((n, 2, 4))), cast(rng.standard_normal((n, qo.OBS_Z))), cast(rng.uniform(-1, 1, (n, qo.RESET_Z)))
        env.set_noise(ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z, obs_z=obs_z, reset_z=rz)
        assert np.array_equal(d.cpu().numpy(), do)
        assert rel_err(o.double().cpu().numpy(), oo) < tol and np.max(np.abs(r.double().cpu().numpy() - ro)) < tol, t
    env.close()


@pytest.mark.parametrize("obs_repr,init", [("xyz_vxyz_R_omega_t2w_t2t", "pitch_roll_restricted"), ("xyzr_vxyzr_quat_omega_h", True),
                                            ("xyzr_vxyzr_R_omega_t2w", False), ("xyzr_vxyzr_quat_omega", False),
                                            ("xyz_vxyz_R_omega_t2w", True)])
def test_fp64_philox_rollout_with_the_new_builders_equals_oracle(obs_repr, init):
    """Philox mode, free running, auto-reset: the t2w / t2t draws (slot 11 step, 41 reset), the quaternion builders and the
    restricted spawn law, against the oracle fed by oracle/philox.py with the same (seed, env id, step, slot) draws."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    n, T, seed, off = 200, 36, 99, 5000
    sn = dict(pos_norm_std=0.005, vel_norm_std=0.01, quat_norm_std=0.01, quat_unif_range=0.02, gyro_noise_density=1.75e-4)
    kw = dict(sim_steps=2, ep_time=0.08, sense_noise=sn, obs_repr=obs_repr, init_random_state=init, t2w_std=0.6, t2t_std=0.5)
    if init == "pitch_roll_restricted":
        kw.update(pitch_max=0.3, roll_max=0.6, yaw_max=1.0)
    env = QuadVecEnv(n, dynamics_params="DefaultQuad", controller="raw", dtype="float64", noise_mode="philox", seed=seed,
                     env_id_offset=off, auto_reset=True, **kw)
    oe = qo.OracleVecEnv(quad_models.defaultquad_params(), n, controller=qo.CTRL_RAW, auto_reset=True, **kw)
    tape = ph.PhiloxTape(seed, fp32=False, sim_steps=2, init_random_state=init)
    ids = np.arange(off, off + n, dtype=np.uint64)
    rng = np.random.RandomState(1)
    s = env.clock[1]
    o = env.reset().cpu().numpy()
    oo = oe.reset(reset_z=tape.reset_z(ids, s))
    assert np.all(np.abs(o - oo) <= obs_tolerance(obs_repr, oo, 1e-11))
    assert o.shape[1] == qo.obs_dim(obs_repr)
    worst = 0.0
    for t in range(T):
        s = env.clock[1]
        a = rng.uniform(-1, 1, (n, 4))
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=tape.ou_z(ids, s), obs_z=tape.obs_z(ids, s), reset_z=tape.reset_z(ids, s))
        assert np.array_equal(d.cpu().numpy(), do)
        e = np.abs(o.cpu().numpy() - oo) / (obs_tolerance(obs_repr, oo, 1.0))
        worst = max(worst, float(e.max()), float(np.max(np.abs(r.cpu().numpy() - ro))))
        assert worst < 1e-8, (t, worst)
    if "t2w" in obs_repr:                                                # wide t2w_std: the clip at [1.5, 10] is exercised
        col = env.obs[:, 18].cpu().numpy()
        assert col.min() >= 0.0 and col.max() <= 1.0
    env.close()


def test_fp32_philox_t2w_and_quat_observations_are_sane():
    """fp32 Philox mode, many envs: t2w / t2t columns are the scaled model constants plus noise of the stated width; the
    quaternion columns are unit and reproduce R."""
    from gym_art_b200 import QuadVecEnv
    n = 100000
    env = QuadVecEnv(n, dynamics_params="DefaultQuad", obs_repr="xyz_vxyz_R_omega_t2w_t2t", seed=5, t2w_std=0.1, t2t_std=0.2)
    env.reset()
    o, _, _ = env.step(env.torch.zeros((n, 4), device="cuda"))
    m = env.model
    t2w, t2t = o[:, 18].double().cpu().numpy(), o[:, 19].double().cpu().numpy()
    w = m.thrust_to_weight
    assert abs(t2w.mean() - (w - 1.5) / 8.5) < 3e-4 and abs(t2w.std() - 0.05 * w / 8.5) < 3e-4
    w = m.torque_to_thrust
    assert abs(t2t.mean() - (w - 0.005) / 0.995) < 3e-5 and abs(t2t.std() - 0.1 * w / 0.995) < 3e-5
    env.close()
    # (yaw_max = 1: with the default 3.14 some attitudes are half turns, where the reference's single-branch R2quat divides by
    # w -> 0 and, in fp32, 1 + trace can round below zero - NaN by the reference's own formula)
    env = QuadVecEnv(n, dynamics_params="DefaultQuad", obs_repr="xyz_vxyz_quat_omega", seed=6, init_random_state="pitch_roll_restricted",
                     yaw_max=1.0)
    env.reset()
    o, _, _ = env.step(env.torch.rand((n, 4), device="cuda") * 2 - 1)
    q = o[:, 6:10].double().cpu().numpy()
    assert o.shape[1] == 13 and np.max(np.abs(np.linalg.norm(q, axis=1) - 1.0)) < 1e-5
    R = env.rot.double().cpu().numpy()
    want = qo.R2quat(R)
    assert np.max(np.abs(q - want)) < 1e-5                               # restricted attitudes: trace well above -1 here
    env.close()

"""GPU: the swarm env step (per-drone env step + pairwise contact in warp shuffles) against the NumPy oracle of the
repo's own swarm specification (no reference semantics exist: parity is oracle parity)."""
import numpy as np
import pytest

from oracle import quadsim_oracle as qo

pytestmark = pytest.mark.gpu


def _pair(E, M, dtype, seed, info=True, formation=False, **kw):
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
    params = quad_models.defaultquad_params()
    goals = np.tile(ring_goals(M, radius=0.3), (E, 1))          # tight ring: contacts happen within a few dozen steps
    base = dict(ep_time=0.4, sim_steps=2, collision_k=1.5)
    base.update(kw)
    coll = base.pop("collision", 0.7)
    oe = qo.OracleVecEnv(params, E * M, controller=qo.CTRL_RAW, auto_reset=True, n_agents=M, goals=goals, box=0.5,
                         collision_coeff=coll, multi_reward=True, **base)
    gk = dict(formation=ring_goals(M, radius=0.3)) if formation else dict(goals=goals)
    env = QuadVecEnv(E * M, dynamics_params=params, controller="raw", dtype=dtype, noise_mode="injected", auto_reset=True,
                     n_agents=M, box=0.5, multi_reward=True, rew_coeff={"collision": coll}, info=info, seed=seed,
                     **gk, **base)
    return oe, env


@pytest.mark.parametrize("M", [2, 8, 32])
def test_swarm_step_matches_oracle_fp64(M):
    E, T = 64 + 5, 100
    rng = np.random.RandomState(M)
    oe, env = _pair(E, M, "float64", 1)
    n = E * M
    rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
    env.set_noise(reset_z=rz)
    np.testing.assert_allclose(env.reset().cpu().numpy(), oe.reset(reset_z=rz), rtol=0, atol=1e-12)
    hits = 0
    for t in range(T):
        a = rng.uniform(-0.2, 0.4, (n, 4))
        ou_z = rng.standard_normal((n, 2, 4))
        rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
        env.set_noise(ou_z=ou_z, reset_z=rz)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z, reset_z=rz)
        assert np.array_equal(d.cpu().numpy(), do)
        # free-running rollout: the device adds a drone's contact impulses in rotated partner order, the oracle in
        # index order, and that last-bit difference grows with the (chaotic) contact dynamics
        np.testing.assert_allclose(o.cpu().numpy(), oo, rtol=1e-8, atol=1e-8, err_msg="obs step %d" % t)
        np.testing.assert_allclose(r.cpu().numpy(), ro, rtol=0, atol=1e-9)
        coll = env.info[:, 17].cpu().numpy() != 0
        assert np.array_equal(coll, oe.collided)
        hits += int(coll.sum())
    assert hits > 0, "the scenario must exercise the contact branch"
    np.testing.assert_allclose(env.stats(), oe.stats, rtol=1e-9)


def test_swarm_step_matches_oracle_fp32_per_step():
    """fp32 build, state re-synchronised every step: 1e-5 relative on velocities / observations after contact."""
    E, M, T = 128, 8, 60
    rng = np.random.RandomState(11)
    oe, env = _pair(E, M, "float32", 2, ep_time=5)
    n = E * M
    rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
    env.set_noise(reset_z=rz)
    env.reset()
    oe.reset(reset_z=rz)
    hits = 0
    for t in range(T):
        for f, v in (("pos", oe.pos), ("vel", oe.vel), ("rot", oe.rot.reshape(n, 9)), ("omega", oe.omega),
                     ("thrust_rot_damp", oe.rot_damp), ("thrust_cmds_damp", oe.cmds_damp), ("ou_state", oe.ou)):
            env.set_field(f, v)
        # the oracle continues from the fp32-rounded state the device holds
        oe.pos, oe.vel, oe.omega = (env.get_field(k).double().cpu().numpy() for k in ("pos", "vel", "omega"))
        oe.rot = env.get_field("rot").double().cpu().numpy()
        oe.rot_damp, oe.cmds_damp, oe.ou = (env.get_field(k).double().cpu().numpy() for k in ("thrust_rot_damp", "thrust_cmds_damp", "ou_state"))
        a = rng.uniform(-0.2, 0.4, (n, 4)).astype(np.float32).astype(np.float64)
        ou_z = rng.standard_normal((n, 2, 4)).astype(np.float32).astype(np.float64)
        env.set_noise(ou_z=ou_z)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z)
        coll = env.info[:, 17].cpu().numpy() != 0
        same = coll == oe.collided                      # a pair within 1 ulp of the contact distance may differ
        assert same.mean() > 0.999
        hits += int(coll.sum())
        ok = np.repeat(same.reshape(E, M).all(axis=1), M)
        err = np.abs(o.double().cpu().numpy() - oo)[ok] / np.maximum(np.abs(oo[ok]), 1.0)
        assert err.max() < 1e-5, (t, err.max())
        assert np.max(np.abs(r.double().cpu().numpy() - ro)[ok]) < 1e-6
    assert hits > 0


@pytest.mark.parametrize("M", [4, 8])
def test_swarm_fast_kernel_matches_oracle_fp32(M):
    """The specialised swarm variant of the TMA kernel (formation goals, no info buffer), per-step against the oracle;
    E is large enough for the multi-tile (mbarrier pipeline) path of the persistent kernel."""
    E, T = (148 * 4 * 128 * 2) // M + 3, 30
    rng = np.random.RandomState(13 + M)
    oe, env = _pair(E, M, "float32", 3, info=False, formation=True, ep_time=5)
    n = E * M
    rz = rng.uniform(-1, 1, (n, qo.RESET_Z))
    env.set_noise(reset_z=rz)
    env.reset()
    oe.reset(reset_z=rz)
    contacts = 0
    for t in range(T):
        for f, v in (("pos", oe.pos), ("vel", oe.vel), ("rot", oe.rot.reshape(n, 9)), ("omega", oe.omega),
                     ("thrust_rot_damp", oe.rot_damp), ("thrust_cmds_damp", oe.cmds_damp), ("ou_state", oe.ou)):
            env.set_field(f, v)
        oe.pos, oe.vel, oe.omega = (env.get_field(k).double().cpu().numpy() for k in ("pos", "vel", "omega"))
        oe.rot = env.get_field("rot").double().cpu().numpy()
        oe.rot_damp, oe.cmds_damp, oe.ou = (env.get_field(k).double().cpu().numpy() for k in ("thrust_rot_damp", "thrust_cmds_damp", "ou_state"))
        a = rng.uniform(-0.2, 0.4, (n, 4)).astype(np.float32).astype(np.float64)
        ou_z = rng.standard_normal((n, 2, 4)).astype(np.float32).astype(np.float64)
        env.set_noise(ou_z=ou_z)
        o, r, d = env.step(a)
        assert env.kernel_variant == "tma-raw-injected"
        oo, ro, do = oe.step(a, ou_z=ou_z)
        contacts += int(oe.collided.sum())
        err = (np.abs(o.double().cpu().numpy() - oo) / np.maximum(np.abs(oo), 1.0)).max(axis=1)
        rerr = np.abs(r.double().cpu().numpy() - ro)
        bad = (err > 1e-5) | (rerr > 1e-6)              # only pairs within an ulp of the contact distance may differ
        assert bad.mean() < 1e-4, (t, bad.sum(), err.max())
    assert contacts > 100


def test_swarm_facade_shapes_and_contact_flags():
    import torch
    from gym_art_b200.quadrotor_multi.quadrotor_multi import QuadrotorEnvMulti
    env = QuadrotorEnvMulti(num_agents=8, num_envs=16, goal_radius=0.2, collision_k=2.0, info=True, seed=3)
    o = env.reset()
    assert tuple(o.shape) == (16, 8, 18)
    seen = False
    for t in range(200):
        o, r, d, info = env.step(torch.zeros((16, 8, 4), device="cuda"))
        seen = seen or bool(info["collided"].any())
    assert tuple(r.shape) == (16, 8) and tuple(d.shape) == (16, 8) and seen
    single = QuadrotorEnvMulti(num_agents=4, seed=1)
    assert tuple(single.reset().shape) == (4, 18)
    env.close(); single.close()


@pytest.mark.parametrize("dtype,tol", [("float64", 1e-8), ("float32", 1e-5)])
def test_swarm_with_per_env_models_matches_oracle(dtype, tol):
    """Dynamics randomisation in swarms (resample_dynamics, quadrotor.py:1030-1056, has no single-drone restriction): every
    swarm of 4 flies its own model (Hummingbird / CrazyFlie / a RandomQuad), the oracle runs the swarms of each model; fp64
    free-running, fp32 per step from the device's state.  The contact distance follows each swarm's own arm."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
    np.random.seed(4)
    zoo = [quad_models.defaultquad_params(), quad_models.crazyflie_params(), qr.RandomQuad().sample()]
    K, M, E, T = len(zoo), 4, 45, 80
    n = E * M
    index = (np.arange(n) // M) % K                               # model of a drone = model of its swarm
    goals = np.tile(ring_goals(M, radius=0.12), (E, 1))
    kw = dict(ep_time=0.3, sim_steps=2, collision_k=2.5, box=0.3, multi_reward=True)
    env = QuadVecEnv(n, dynamics_params=zoo[0], controller="raw", dtype=dtype, noise_mode="injected", auto_reset=True, n_agents=M,
                     goals=goals, rew_coeff={"collision": 0.7}, info=True, per_env_dynamics=K, **kw)
    env.set_models(zoo, index=index)
    sel = [np.where(index == k)[0] for k in range(K)]
    oes = [qo.OracleVecEnv(zoo[k], len(sel[k]), controller=qo.CTRL_RAW, auto_reset=True, n_agents=M, goals=goals[sel[k]],
                           collision_coeff=0.7, **kw) for k in range(K)]
    rng = np.random.RandomState(21)
    cast = (lambda x: x.astype(np.float32).astype(np.float64)) if dtype == "float32" else (lambda x: x)
    rz = cast(rng.uniform(-1, 1, (n, qo.RESET_Z)))
    env.set_noise(reset_z=rz)
    o = env.reset().double().cpu().numpy()
    for k in range(K):
        np.testing.assert_allclose(o[sel[k]], oes[k].reset(reset_z=rz[sel[k]]), rtol=0, atol=max(tol, 1e-12))
    fields = (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
              ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou"))
    hits = 0
    for t in range(T):
        if dtype == "float32":
            dev = {f: env.get_field(f).double().cpu().numpy() for f, _ in fields}
            for k in range(K):
                for f, of in fields:
                    setattr(oes[k], of, dev[f][sel[k]].copy())
        a, ou_z, rz = cast(rng.uniform(-0.2, 0.4, (n, 4))), cast(rng.standard_normal((n, 2, 4))), cast(rng.uniform(-1, 1, (n, qo.RESET_Z)))
        env.set_noise(ou_z=ou_z, reset_z=rz)
        o, r, d = env.step(a)
        o, r, d = o.double().cpu().numpy(), r.double().cpu().numpy(), d.cpu().numpy()
        coll = env.info[:, 17].cpu().numpy() != 0
        for k in range(K):
            s = sel[k]
            oo, ro, do = oes[k].step(a[s], ou_z=ou_z[s], reset_z=rz[s])
            assert np.array_equal(d[s], do)
            same = coll[s] == oes[k].collided                     # fp32: a pair within an ulp of the contact distance may differ
            assert same.all() if dtype == "float64" else same.mean() > 0.99
            ok = np.repeat(same.reshape(-1, M).all(axis=1), M)
            assert np.max(np.abs(o[s] - oo)[ok] / np.maximum(np.abs(oo[ok]), 1.0)) < tol, (k, t)
            assert np.max(np.abs(r[s] - ro)[ok]) < max(tol * 0.1, 1e-9)
        hits += int(coll.sum())
    assert hits > 0, "the scenario must exercise the contact branch"
    assert env.kernel_variant == ("generic-permodel-injected" if dtype == "float64" else "tma-generic-permodel-injected")
    env.close()


def test_mixed_model_swarm_contact_distance_is_the_sum_of_the_arms():
    """Known answer for drones of different size in one swarm: the pair (i, j) touches iff d < collision_k (arm_i + arm_j).
    Two drones 0.2 m apart on x, closing at 1 m/s each: arms 0.17 (Hummingbird) and 0.046 (CrazyFlie).  collision_k = 1:
    0.216 > 0.2 -> contact, equal-mass exchange swaps the velocities; collision_k = 0.9: 0.194 < 0.2 -> nothing happens."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    zoo = [quad_models.defaultquad_params(), quad_models.crazyflie_params()]
    arms = [qo.model_constants(z)["arm"] for z in zoo]
    assert 0.9 * sum(arms) < 0.2 < sum(arms)
    for k, swapped in ((1.0, True), (0.9, False)):
        env = QuadVecEnv(8, dynamics_params=zoo[0], controller="raw", dtype="float64", noise_mode="injected", auto_reset=False,
                         n_agents=2, sim_freq=1e9, sim_steps=1, ep_time=1e-6, collision_k=k, info=True, per_env_dynamics=2,
                         goals=np.tile([0.0, 0.0, 2.0], (8, 1)))
        env.set_models(zoo, index=np.arange(8) % 2)
        pos = np.tile(np.array([[0.0, 0.0, 2.0], [0.2, 0.0, 2.0]]), (4, 1))
        vel = np.tile(np.array([[1.0, 0.0, 0.0], [-1.0, 0.0, 0.0]]), (4, 1))
        env.set_state(pos, vel, np.tile(np.eye(3), (8, 1, 1)), np.zeros((8, 3)))
        env.step(-np.ones((8, 4)))
        v = env.vel.cpu().numpy()
        want = -vel if swapped else vel
        np.testing.assert_allclose(v[:, 0], want[:, 0], rtol=0, atol=1e-6)
        assert ((env.info[:, 17].cpu().numpy() != 0) == swapped).all()
        # the oracle's statement of the same rule
        oe = qo.OracleVecEnv(zoo[0], 2, n_agents=2, auto_reset=False, collision_k=k)
        oe.arm_env = np.array(arms)
        oe.pos, oe.vel = pos[:2].copy(), vel[:2].copy()
        oe.collide_swarm()
        np.testing.assert_allclose(oe.vel, want[:2], rtol=0, atol=1e-12)
        env.close()


def test_swarm_with_rotor_drag_takes_a_kernel_that_has_the_contact_stage():
    """A swarm whose model needs another feature group (rotor drag here) must not land in that group's single-feature
    kernel, which has no contact stage: fp32 per step against the oracle, contacts flagged."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
    params = quad_models.crazyflie_params()
    qr._update_existing(params, {"motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424}, "damp": {"vel": 0.001, "omega_quadratic": 0.015}})
    E, M, T = 300, 4, 40
    n = E * M
    form = ring_goals(M, radius=0.05)
    kw = dict(ep_time=5, sim_steps=2, collision_k=2.5, box=0.15)
    env = QuadVecEnv(n, dynamics_params=params, controller="raw", noise_mode="injected", auto_reset=True, n_agents=M, formation=form,
                     rew_coeff={"collision": 0.7}, **kw)
    oe = qo.OracleVecEnv(params, n, controller=qo.CTRL_RAW, auto_reset=True, n_agents=M, goals=np.tile(form, (E, 1)), collision_coeff=0.7, **kw)
    rng = np.random.RandomState(31)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    rz = f32(rng.uniform(-1, 1, (n, qo.RESET_Z)))
    env.set_noise(reset_z=rz)
    env.reset()
    oe.reset(reset_z=rz)
    fields = (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
              ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou"))
    contacts = 0
    for t in range(T):
        for f, of in fields:
            setattr(oe, of, env.get_field(f).double().cpu().numpy())
        a, ou_z = f32(rng.uniform(-0.2, 0.4, (n, 4))), f32(rng.standard_normal((n, 2, 4)))
        env.set_noise(ou_z=ou_z)
        o, r, d = env.step(a)
        oo, ro, do = oe.step(a, ou_z=ou_z)
        contacts += int(oe.collided.sum())
        err = (np.abs(o.double().cpu().numpy() - oo) / np.maximum(np.abs(oo), 1.0)).max(axis=1)
        bad = (err > 1e-5) | (np.abs(r.double().cpu().numpy() - ro) > 1e-6)   # only pairs within an ulp of the contact distance may differ
        assert bad.mean() < 2e-3, (t, bad.sum(), err.max())
    assert contacts > 50
    assert env.kernel_variant == "tma-generic-injected"                       # drag + the swarm stage: two groups, the catch-all
    env.close()


def test_swarm_with_per_env_models_specialised_twin_fp32():
    """The same scenario without the info buffer: per-env goals + per-env models + RawControl take the compile-time-controller
    twin of the F_MISC group with per-env models (`tma-raw`), whose contact stage reads each partner's arm by shuffle."""
    from gym_art_b200 import QuadVecEnv
    from gym_art_b200.quadrotor import quad_models
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
    np.random.seed(4)
    zoo = [quad_models.defaultquad_params(), quad_models.crazyflie_params(), qr.RandomQuad().sample()]
    K, M, T = len(zoo), 4, 40
    E = (148 * 3 * 128 * 2) // M + 3                              # enough tiles for the pipelined (multi-tile) path
    n = E * M
    index = (np.arange(n) // M) % K
    goals = np.tile(ring_goals(M, radius=0.12), (E, 1))
    kw = dict(ep_time=5, sim_steps=2, collision_k=2.5, box=0.3, multi_reward=True)
    env = QuadVecEnv(n, dynamics_params=zoo[0], controller="raw", noise_mode="injected", auto_reset=True, n_agents=M, goals=goals,
                     rew_coeff={"collision": 0.7}, per_env_dynamics=K, **kw)
    env.set_models(zoo, index=index)
    sel = [np.where(index == k)[0] for k in range(K)]
    oes = [qo.OracleVecEnv(zoo[k], len(sel[k]), controller=qo.CTRL_RAW, auto_reset=True, n_agents=M, goals=goals[sel[k]],
                           collision_coeff=0.7, **kw) for k in range(K)]
    rng = np.random.RandomState(22)
    f32 = lambda x: x.astype(np.float32).astype(np.float64)
    rz = f32(rng.uniform(-1, 1, (n, qo.RESET_Z)))
    env.set_noise(reset_z=rz)
    env.reset()
    for k in range(K):
        oes[k].reset(reset_z=rz[sel[k]])
    fields = (("pos", "pos"), ("vel", "vel"), ("rot", "rot"), ("omega", "omega"), ("thrust_rot_damp", "rot_damp"),
              ("thrust_cmds_damp", "cmds_damp"), ("ou_state", "ou"))
    contacts = 0
    for t in range(T):
        dev = {f: env.get_field(f).double().cpu().numpy() for f, _ in fields}
        for k in range(K):
            for f, of in fields:
                setattr(oes[k], of, dev[f][sel[k]].copy())
        a, ou_z = f32(rng.uniform(-0.2, 0.4, (n, 4))), f32(rng.standard_normal((n, 2, 4)))
        env.set_noise(ou_z=ou_z)
        o, r, d = env.step(a)
        o, r = o.double().cpu().numpy(), r.double().cpu().numpy()
        for k in range(K):
            s = sel[k]
            oo, ro, do = oes[k].step(a[s], ou_z=ou_z[s])
            contacts += int(oes[k].collided.sum())
            err = (np.abs(o[s] - oo) / np.maximum(np.abs(oo), 1.0)).max(axis=1)
            bad = (err > 1e-5) | (np.abs(r[s] - ro) > 1e-6)       # only pairs within an ulp of the contact distance may differ
            assert bad.mean() < 1e-3, (k, t, bad.sum(), err.max())
    assert contacts > 100
    assert env.kernel_variant == "tma-raw-injected"
    env.close()

"""CPU: properties of the swarm-contact specification (oracle collide_swarm).  The reference has no multi-drone
semantics (parity unpinned), so these pin the spec itself: what the CUDA path is then
held to on the GPU (tests/test_gpu_swarm.py)."""
import numpy as np

from oracle import quadsim_oracle as qo
from gym_art_b200.quadrotor import quad_models


def _swarm(E, M, seed=0, spread=0.25, **kw):
    rng = np.random.RandomState(seed)
    env = qo.OracleVecEnv(quad_models.defaultquad_params(), E * M, n_agents=M, auto_reset=False, **kw)
    env.pos = rng.uniform(-spread, spread, (E * M, 3)) + np.array([0, 0, 2.0])
    env.vel = rng.uniform(-1, 1, (E * M, 3))
    return env


def test_pair_exchange_conserves_momentum_and_energy():
    """Two equal masses: the normal velocity components are swapped -> momentum and kinetic energy are conserved."""
    env = _swarm(200, 2, seed=1, spread=0.3)
    v0 = env.vel.copy()
    env.collide_swarm()
    v1 = env.vel
    hit = env.collided.reshape(-1, 2)
    assert hit.any() and not hit.all()
    assert np.array_equal(hit[:, 0], hit[:, 1])                     # contact is symmetric
    np.testing.assert_allclose(v1.reshape(-1, 2, 3).sum(1), v0.reshape(-1, 2, 3).sum(1), atol=1e-12)
    np.testing.assert_allclose((v1 ** 2).reshape(-1, 2, 3).sum((1, 2)), (v0 ** 2).reshape(-1, 2, 3).sum((1, 2)), atol=1e-12)
    # separating pairs in contact are left alone
    changed = np.any(v1 != v0, axis=1).reshape(-1, 2)
    assert np.array_equal(changed[:, 0], changed[:, 1]) and (hit[:, 0] & ~changed[:, 0]).any()


def test_after_exchange_pairs_separate():
    env = _swarm(300, 2, seed=2, spread=0.1)
    p = env.pos.reshape(-1, 2, 3)
    env.collide_swarm()
    v = env.vel.reshape(-1, 2, 3)
    n = p[:, 0] - p[:, 1]
    closing = np.sum((v[:, 1] - v[:, 0]) * n, axis=1)               # > 0 means still approaching
    hit = env.collided.reshape(-1, 2)[:, 0]
    assert np.all(closing[hit] <= 1e-12)


def test_order_independent_and_groups_do_not_interact():
    """Relabelling the drones inside a swarm permutes the result; drones of different swarms never touch."""
    E, M = 50, 8
    env = _swarm(E, M, seed=3)
    perm = np.random.RandomState(4).permutation(M)
    env2 = _swarm(E, M, seed=3)
    env2.pos = env.pos.reshape(E, M, 3)[:, perm].reshape(E * M, 3)
    env2.vel = env.vel.reshape(E, M, 3)[:, perm].reshape(E * M, 3)
    env.collide_swarm()
    env2.collide_swarm()
    np.testing.assert_allclose(env2.vel.reshape(E, M, 3), env.vel.reshape(E, M, 3)[:, perm], atol=1e-12)
    assert np.array_equal(env2.collided.reshape(E, M), env.collided.reshape(E, M)[:, perm])
    # all swarms stacked on the same spot: a single-swarm run of swarm 0 gives the same answer as inside the batch
    one = _swarm(1, M, seed=3)
    one.pos, one.vel = _swarm(E, M, seed=3).pos[:M].copy(), _swarm(E, M, seed=3).vel[:M].copy()
    one.collide_swarm()
    np.testing.assert_allclose(one.vel, env.vel[:M], atol=0)


def test_far_apart_is_identity_and_collision_cost_enters_reward():
    E, M = 4, 4
    env = _swarm(E, M, seed=5, spread=3.0, collision_coeff=2.0, rew_coeff={"collision": 2.0} if False else None)
    v0 = env.vel.copy()
    env.collide_swarm()
    if not env.collided.any():
        assert np.array_equal(env.vel, v0)
    # reward: a drone in contact pays collision_coeff * dt on top of the single-drone reward
    a = _swarm(1, 2, seed=6, spread=0.05, collision_coeff=3.0)
    b = _swarm(1, 2, seed=6, spread=0.05, collision_coeff=0.0)
    act = np.zeros((2, 4))
    _, ra, _ = a.step(act)
    _, rb, _ = b.step(act)
    assert a.collided.all()
    np.testing.assert_allclose(rb - ra, 3.0 * a.dt, atol=1e-12)

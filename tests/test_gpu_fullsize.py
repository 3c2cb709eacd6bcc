"""GPU: the BASELINE.json configurations at their FULL sizes, checked through size-independent properties (the
oracle cannot run millions of envs): invariance to batch size / position in the batch (Philox is keyed by the global
env id, so any window of a huge run must equal a small run over the same ids bit for bit), rotation matrices stay
orthonormal, states stay inside the room, episode bookkeeping is exact, nothing is non-finite."""
import pytest

pytestmark = pytest.mark.gpu


def _cfg(name):
    from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
    base = dict(ep_time=0.1, seed=11)            # 11-step episodes: several auto-resets inside a short run
    if name == "cfg2":
        return 65536, dict(dynamics_params="DefaultQuad", controller="raw", **base)
    if name == "cfg3":
        return 1 << 20, dict(dynamics_params="Crazyflie", controller="omega_thrust", sense_noise="default", init_random_state=True, **base)
    if name == "cfg4":
        return 262144 * 8, dict(dynamics_params="DefaultQuad", controller="raw", n_agents=8, formation=ring_goals(8), **base)
    if name == "cfg5":
        return 8 << 20, dict(dynamics_params="DefaultQuad", controller="mellinger", **base)     # 64 M envs / 8 GPUs
    if name == "cfg5_64M":
        return 64 << 20, dict(dynamics_params="DefaultQuad", controller="mellinger", **base)    # the whole job on one GPU
    raise ValueError(name)


def _actions(torch, n, act_dim, offset, T, cfg):
    """Actions as a pure function of (global env id, t), so a window can regenerate exactly its slice."""
    if cfg.startswith("cfg5"):
        return [None] * T
    ids = torch.arange(offset, offset + n, device="cuda", dtype=torch.float64)
    out = []
    for t in range(T):
        ph = ids[:, None] * torch.tensor([0.137, 0.291, 0.403, 0.577], device="cuda", dtype=torch.float64) + 0.31 * t
        a = torch.sin(ph).float()
        if cfg == "cfg3":
            a[:, 0] *= 0.2
        out.append(a.contiguous())
    return out


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3", "cfg4", "cfg5", "cfg5_64M"])
def test_full_size_run_equals_windows_and_keeps_invariants(cfg):
    import torch
    from gym_art_b200 import QuadVecEnv
    n, kw = _cfg(cfg)
    T, W = (25, 4096) if n < (32 << 20) else (13, 4096)      # 64 M envs: one whole 11-step episode + a reset
    env = QuadVecEnv(n, **kw)
    env.reset()
    acts = _actions(torch, n, env.act_dim, 0, T, cfg)
    dones = torch.zeros(n, device="cuda", dtype=torch.int32)
    for t in range(T):
        o, r, d = env.step(acts[t])
        dones += d.int()
    env.check_finite()
    assert bool(torch.isfinite(o).all()) and bool(torch.isfinite(r).all())
    # ---- invariants over ALL envs
    R = env.rot
    err = (R.transpose(1, 2) @ R - torch.eye(3, device="cuda")).abs().amax()
    assert float(err) < 2e-5, "rotation matrices drifted from orthonormal: %g" % float(err)
    pos = env.pos
    assert float(pos[:, :2].abs().max()) <= 10.0 and float(pos[:, 2].min()) >= 0.0 and float(pos[:, 2].max()) <= 10.0
    assert float(env.omega.abs().max()) <= 40.0 + 1e-4
    cd = env.thrust_cmds_damp
    assert float(cd.min()) >= 0.0 and float(cd.max()) <= 1.0
    ep = env.ep_len + 1                                           # episodes last ep_len + 1 steps (quadrotor.py:986-987)
    assert int(dones.min()) == int(dones.max()) == T // ep
    assert bool((env.tick == T % ep).all())
    st = env.stats()
    assert st[2] == float(n) * (T // ep)
    # ---- any window of the big run == a small run over the same global env ids (bit for bit)
    full_obs, full_rew, full_state = o.clone(), r.clone(), env.state.clone()
    env.close()
    del env
    m = kw.get("n_agents", 1)
    for offset in (0, (n // 3 // 128) * 128 + 128 * 5, n - W):
        offset -= offset % m
        small = QuadVecEnv(W, env_id_offset=offset, **kw)
        small.reset()
        a_s = _actions(torch, W, small.act_dim, offset, T, cfg)
        for t in range(T):
            so, sr, sd = small.step(a_s[t])
        assert torch.equal(so, full_obs[offset:offset + W]), (cfg, offset)
        assert torch.equal(sr, full_rew[offset:offset + W])
        assert torch.equal(small.state, full_state[:, offset:offset + W])
        small.close()

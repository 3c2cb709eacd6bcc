"""quadrotor_multi on the B200 path.

`QuadrotorEnv` mirrors gym_art/quadrotor_multi/quadrotor_multi.py - which, despite its name, is a single-drone fork
of quadrotor.py (SURVEY.md "five facts" #1): log-distance position reward `pos * (log_w * log(dist + off) + lin_w *
dist)` (:554-555), defaults effort 0.01 / spin 0 (:810-818), ep_time 4 (:668), sampler dicts keyed by "type" (:759).

`QuadrotorEnvMulti` is the swarm env north_star asks for.  The reference has no multi-drone state, collision test
or response anywhere, so its semantics are this repo's own specification (DESIGN.md "Swarm contact", restated in
oracle/quadsim_oracle.py collide_swarm): `num_agents` drones share one room, each flies the single-drone env step
towards its own goal; after the substeps of an env step every pair closer than collision_k * 2 * arm exchanges the
approaching normal velocity component (equal masses, evaluated from pre-contact velocities, order independent), and
a drone in contact pays `rew_coeff["collision"]` per step.  One warp-aligned group of lanes per swarm on the GPU.
"""
import numpy as np

from .. import _lib
from ..quadrotor.quadrotor import GRAV, QuadrotorEnv as _SingleEnv
from ..vec_env import QuadVecEnv

_ALIASES = {"defaultquad": "DefaultQuad", "crazyflie": "Crazyflie", "mediumquad": "MediumQuad"}


class QuadrotorEnv(_SingleEnv):
    """Single drone, quadrotor_multi's reward variant and defaults (quadrotor_multi.py:660-671, :805-830)."""
    _multi_reward = True
    _sampler_key = "type"

    def __init__(self, dynamics_params="DefaultQuad", dynamics_change=None, dynamics_randomize_every=None,
                 dyn_sampler_1=None, dyn_sampler_2=None, raw_control=True, raw_control_zero_middle=True, dim_mode='3D',
                 tf_control=False, sim_freq=200., sim_steps=2, obs_repr="xyz_vxyz_R_omega", ep_time=4, obstacles_num=0,
                 room_size=10, init_random_state=False, rew_coeff=None, sense_noise=None, verbose=False, gravity=GRAV,
                 resample_goal=False, t2w_std=0.005, t2t_std=0.0005, excite=False, dynamics_simplification=False, **device_kw):
        # the reference's own default "defaultquad" names no sampler class and raises AttributeError; accept it
        if isinstance(dynamics_params, str):
            dynamics_params = _ALIASES.get(dynamics_params, dynamics_params)
        _SingleEnv.__init__(self, dynamics_params, dynamics_change, dynamics_randomize_every, dyn_sampler_1, dyn_sampler_2,
                            raw_control, raw_control_zero_middle, dim_mode, tf_control, sim_freq, sim_steps, obs_repr,
                            ep_time, obstacles_num, room_size, init_random_state, rew_coeff, sense_noise, verbose, gravity,
                            resample_goal, t2w_std, t2t_std, excite, dynamics_simplification, **device_kw)


def ring_goals(num_agents, radius=1.0, height=2.0):
    """Default swarm formation: goals evenly spaced on a horizontal ring (BASELINE.json configs[3])."""
    ang = 2 * np.pi * np.arange(num_agents) / num_agents
    return np.stack([radius * np.cos(ang), radius * np.sin(ang), np.full(num_agents, height)], axis=1)


class QuadrotorEnvMulti(object):
    """`num_envs` swarms of `num_agents` drones.  step(actions [E, M, A]) -> (obs [E, M, O], rew [E, M], done [E, M],
    info) as torch tensors on the device (E = 1 is squeezed away when `num_envs` is None, giving the per-agent lists
    multi-agent RL code expects)."""

    def __init__(self, num_agents=8, num_envs=None, dynamics_params="DefaultQuad", raw_control=True,
                 raw_control_zero_middle=True, sim_freq=200., sim_steps=2, obs_repr="xyz_vxyz_R_omega", ep_time=4,
                 room_size=10, init_random_state=False, rew_coeff=None, sense_noise=None, collision_k=1.0, goals=None,
                 goal_radius=1.0, multi_reward=True, device="cuda:0", dtype="float32", seed=0, noise_mode="philox",
                 auto_reset=True, info=False):
        self.num_agents = int(num_agents)
        self._squeeze = num_envs is None
        self.num_envs = 1 if num_envs is None else int(num_envs)
        E, M = self.num_envs, self.num_agents
        if goals is None:
            goals = ring_goals(M, goal_radius)
        goals = np.asarray(goals, dtype=np.float64)
        # one formation shared by every swarm lives in the kernel's constant block; per-swarm goals need a state row
        formation = goals.reshape(M, 3) if goals.size == M * 3 else None
        goals = np.broadcast_to(goals.reshape(-1, M, 3), (E, M, 3)).reshape(E * M, 3).copy()
        if isinstance(dynamics_params, str):
            dynamics_params = _ALIASES.get(dynamics_params, dynamics_params)
        self.vec = QuadVecEnv(E * M, dynamics_params=dynamics_params, raw_control=raw_control,
                              raw_control_zero_middle=raw_control_zero_middle, sim_freq=sim_freq, sim_steps=sim_steps,
                              obs_repr=obs_repr, ep_time=ep_time, room_size=room_size, init_random_state=init_random_state,
                              rew_coeff=rew_coeff, sense_noise=sense_noise, device=device, dtype=dtype, seed=seed,
                              noise_mode=noise_mode, auto_reset=auto_reset, n_agents=M,
                              goals=None if formation is not None else goals, formation=formation,
                              multi_reward=multi_reward, collision_k=collision_k, info=info)
        self.goals = goals.reshape(E, M, 3)
        self.obs_dim, self.act_dim, self.ep_len, self.dt = self.vec.obs_dim, self.vec.act_dim, self.vec.ep_len, self.vec.dt
        self.rew_coeff = self.vec.rew_coeff

    def _shape(self, t, *tail):
        E, M = self.num_envs, self.num_agents
        return t.reshape(*(((M,) if self._squeeze else (E, M)) + tail))

    def reset(self):
        return self._shape(self.vec.reset(), self.obs_dim)

    def step(self, actions=None):
        a = None
        if actions is not None:
            a = self.vec.torch.as_tensor(actions, dtype=self.vec.tdtype, device=self.vec.device).reshape(
                self.num_envs * self.num_agents, self.act_dim)
        obs, rew, done = self.vec.step(a)
        info = {}
        if self.vec.info is not None:
            info["collided"] = self._shape(self.vec.info[:, _lib.QS_INFO_COLLIDED] != 0)
            info["crashed"] = self._shape(self.vec.info[:, _lib.QS_INFO_CRASHED] != 0)
        return self._shape(obs, self.obs_dim), self._shape(rew), self._shape(done), info

    def close(self):
        self.vec.close()

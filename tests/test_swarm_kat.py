"""Hand-derived known answers for the swarm contact stage (tests/golden/swarm_kat.json; derivations in its "why" strings):
an independent pin for the stage the reference does not have.  Both implementations must hit
them - the NumPy oracle here on the CPU, the CUDA kernel in the GPU half of this file."""
import json
import os

import numpy as np
import pytest

from oracle import quadsim_oracle as qo

KAT = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "swarm_kat.json")))
CASES = KAT["cases"]


def _params():
    from gym_art_b200.quadrotor import quad_models
    return quad_models.defaultquad_params()


def _k_for_unit_threshold():
    arm = qo.model_constants(_params())["arm"]
    return 1.0 / (2.0 * arm)                       # contact iff d < collision_k * 2 * arm = 1


@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_oracle_hits_the_hand_derived_answers(case):
    M = case["M"]
    env = qo.OracleVecEnv(_params(), M, n_agents=M, auto_reset=False, collision_k=_k_for_unit_threshold())
    env.pos, env.vel = np.array(case["pos"], dtype=np.float64), np.array(case["vel"], dtype=np.float64)
    env.collide_swarm()
    np.testing.assert_allclose(env.vel, np.array(case["vel_after"]), rtol=0, atol=1e-12)
    assert env.collided.astype(int).tolist() == case["hit"]


def test_closed_form_of_the_32_drone_case():
    c = [x for x in CASES if x["name"] == "thirty_two_all_pairs"][0]
    v = np.array(c["vel_after"])[:, 0]
    assert abs(v[0] + 49.6) < 1e-12 and abs(v[31] - 46.5) < 1e-12
    assert abs(v.sum() - np.array(c["vel"])[:, 0].sum()) < 1e-9          # pairwise equal and opposite impulses: momentum kept


@pytest.mark.gpu
@pytest.mark.parametrize("dtype,tol", [("float64", 1e-6), ("float32", 2e-5)])
@pytest.mark.parametrize("case", CASES, ids=[c["name"] for c in CASES])
def test_kernel_hits_the_hand_derived_answers(case, dtype, tol):
    """One env step with a vanishing time step (dt = 1e-9 s: the dynamics move the state by ~1e-8) so that what comes
    out is the contact stage applied to the given positions and velocities; several copies of the swarm fill the warp
    (M lanes per swarm, 32 / M swarms per warp)."""
    import torch
    from gym_art_b200 import QuadVecEnv
    M, E = case["M"], 5
    env = QuadVecEnv(E * M, dynamics_params=_params(), controller="raw", dtype=dtype, noise_mode="injected", auto_reset=False,
                     n_agents=M, sim_freq=1e9, sim_steps=1, ep_time=1e-6, collision_k=_k_for_unit_threshold(), info=True,
                     room_size=10, goals=np.tile([0.0, 0.0, 2.0], (E * M, 1)))
    pos, vel = np.tile(np.array(case["pos"], dtype=np.float64), (E, 1)), np.tile(np.array(case["vel"], dtype=np.float64), (E, 1))
    env.set_state(pos, vel, np.tile(np.eye(3), (E * M, 1, 1)), np.zeros((E * M, 3)))
    env.step(-np.ones((E * M, 4)))                                       # zero thrust
    v = env.vel.double().cpu().numpy()
    t = max(tol, case.get("tol", 0.0), case.get("tol32", 0.0) if dtype == "float32" else 0.0)
    np.testing.assert_allclose(v, np.tile(np.array(case["vel_after"]), (E, 1)), rtol=0, atol=t * max(1.0, float(np.abs(case["vel_after"]).max())))
    hit = (env.info[:, 17].cpu().numpy() != 0).astype(int)
    assert hit.tolist() == case["hit"] * E
    env.close()

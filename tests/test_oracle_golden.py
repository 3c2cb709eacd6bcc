"""CPU: the NumPy oracle must reproduce the fixtures recorded from the live reference.

Tolerance: the oracle is a float64 restatement with the same operation order; differences come
only from batched-vs-scalar NumPy kernels (BLAS dot vs explicit sums), a few ulp per step.
"""
import json

import numpy as np
import pytest

from oracle import quadsim_oracle as qo
from tests.golden_util import Golden, ROLLOUT_CASES, load_npz, obs_tolerance

ATOL = 2e-12


def make_oracle(g, n=1, auto_reset=False):
    oe = qo.OracleVecEnv(g.params, n, auto_reset=auto_reset, **g.oracle_kwargs)
    oe.since_last_svd = float(g.svd_timer0)
    oe.ou = np.tile(g.ou0, (n, 1))
    oe.gyro_bias = np.tile(g.gyro_bias0, (n, 1))
    return oe


@pytest.mark.parametrize("name", ROLLOUT_CASES)
def test_oracle_reproduces_reference_rollout(name):
    g = Golden(name)
    oe = make_oracle(g)
    A = qo.CTRL_ACT_DIM[g.controller]
    obs0 = oe.reset(reset_z=g.reset_z[0][None], goal_z=g.reset_goal_z[0][None])
    np.testing.assert_allclose(obs0[0], g.reset_obs[0], rtol=0, atol=ATOL)
    for t in range(g.T):
        o, r, d = oe.step(g.actions[t][None, :A], ou_z=g.ou_z[t][None], obs_z=g.obs_z[t][None], goal_z=g.goal_z[t][None])
        assert bool(d[0]) == bool(g.done[t])
        assert np.all(np.abs(o[0] - g.obs[t]) <= obs_tolerance(g.oracle_kwargs["obs_repr"], g.obs[t], ATOL)), "obs step %d" % t
        np.testing.assert_allclose(r[0], g.rew[t], rtol=0, atol=ATOL, err_msg="rew step %d" % t)
        for k in ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "ou", "accelerometer"):
            np.testing.assert_allclose(getattr(oe, k)[0], g.a[k][t], rtol=0, atol=ATOL, err_msg="%s step %d" % (k, t))
        if d[0]:
            i = g.reset_index_after(t)
            ob = oe.reset(reset_z=g.reset_z[i][None], goal_z=g.reset_goal_z[i][None])
            np.testing.assert_allclose(ob[0], g.reset_obs[i], rtol=0, atol=ATOL)


def test_oracle_auto_reset_matches_explicit_reset():
    g = Golden("mq_raw_simsteps4")
    oe = make_oracle(g, auto_reset=True)
    oe.reset(reset_z=g.reset_z[0][None])
    n_done = 0
    for t in range(g.T):
        rz = None
        if g.done[t]:
            rz = g.reset_z[g.reset_index_after(t)][None]
        o, r, d = oe.step(g.actions[t][None], ou_z=g.ou_z[t][None], obs_z=g.obs_z[t][None], reset_z=rz)
        if d[0]:
            n_done += 1
            np.testing.assert_allclose(o[0], g.reset_obs[g.reset_index_after(t)], rtol=0, atol=ATOL)
        else:
            np.testing.assert_allclose(o[0], g.obs[t], rtol=0, atol=ATOL)
        np.testing.assert_allclose(r[0], g.rew[t], rtol=0, atol=ATOL)
    assert n_done == 5 and oe.stats[2] == 5


def test_oracle_known_answer():
    ka = load_npz("known_answer")
    for tag in ("hb", "cf"):
        params = json.loads(str(ka["%s_params_json" % tag]))
        oe = qo.OracleVecEnv(params, 1, ep_time=5, auto_reset=False)
        oe.set_state([0.5, -0.25, 1.5], [0.1, 0.2, -0.3], ka["R0"], [0.5, -0.4, 0.3])
        a = np.array([[0.1, -0.2, 0.3, 0.0]])
        for t in range(100):
            o, r, d = oe.step(a)
            if t + 1 in (1, 100):
                for k in ("pos", "vel", "rot", "omega", "cmds_damp"):
                    np.testing.assert_allclose(getattr(oe, k)[0], ka["%s_%d_%s" % (tag, t + 1, k)], rtol=0, atol=ATOL)
                np.testing.assert_allclose(r[0], ka["%s_%d_rew" % (tag, t + 1)], rtol=0, atol=ATOL)
                assert bool(oe.crashed[0]) == bool(ka["%s_%d_crashed" % (tag, t + 1)])
    # the values transcribed in SURVEY.md section 8c (smoke check of the fixture itself)
    np.testing.assert_allclose(ka["hb_1_rew"], 6.658377083393288e-4, rtol=1e-12)
    np.testing.assert_allclose(ka["hb_100_omega"], [-16.460095973639216, 1.2687078456332608, -13.368468021920048],
                               rtol=1e-12)
    np.testing.assert_allclose(ka["cf_1_cmds_damp"][0], 0.03406970164011542, rtol=1e-12)


def test_oracle_model_constants():
    ma = load_npz("model_answers")
    for name in ("DefaultQuad", "Crazyflie", "MediumQuad"):
        m = qo.model_constants(json.loads(str(ma["%s_params_json" % name])))
        for k, ok in (("mass", "mass"), ("inertia", "inertia"), ("prop_pos", "prop_pos"), ("thrust_max", "thrust_max"),
                      ("torque_max", "torque_max"), ("arm", "arm"), ("prop_crossproducts", "prop_cross"),
                      ("Jinv", "Jinv")):
            np.testing.assert_allclose(m[ok], ma["%s_%s" % (name, k)], rtol=1e-13, atol=1e-18, err_msg=name + k)
    # SURVEY.md section 8 a0 known answers
    np.testing.assert_allclose(ma["DefaultQuad_mass"], 0.816, rtol=1e-12)
    np.testing.assert_allclose(ma["DefaultQuad_inertia"], [3.7465748e-3, 3.7465748e-3, 6.1493417e-3], rtol=1e-7)
    np.testing.assert_allclose(ma["Crazyflie_thrust_max"], 0.130473, rtol=1e-6)
    np.testing.assert_allclose(ma["DefaultQuad_Jinv"][0], [0.0364060, -0.00139295, -0.00139295, -0.00548708], rtol=2e-6)

"""CPU: the product's host-side parameter pipeline (quad_models -> QuadLink -> ModelConstants) against the
constants recorded from the live reference (tests/golden/model_answers.npz) and against
the oracle's independent restatement."""
import json

import numpy as np
import pytest

from gym_art_b200.quadrotor import quad_models
from gym_art_b200.quadrotor.inertia import QuadLink
from gym_art_b200.quadrotor.model import ModelConstants
from gym_art_b200.vec_env import make_config, resolve_dynamics_params
from oracle import quadsim_oracle as qo
from tests.golden_util import load_npz


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, (list, tuple, np.ndarray)):
        return [_plain(v) for v in np.asarray(x).tolist()] if np.ndim(x) else float(x)
    return float(x) if isinstance(x, (int, float, np.floating, np.integer)) else x


@pytest.mark.parametrize("name", ["DefaultQuad", "Crazyflie", "MediumQuad"])
def test_model_constants_match_reference(name):
    ma = load_npz("model_answers")
    ref_params = json.loads(str(ma["%s_params_json" % name]))
    params = resolve_dynamics_params(name)
    assert _plain(params) == _plain(ref_params), "parameter set differs from the reference's"
    m = ModelConstants(params)
    for key, attr in (("mass", "mass"), ("inertia", "inertia"), ("prop_pos", "prop_pos"), ("thrust_max", "thrust_max"),
                      ("torque_max", "torque_max"), ("arm", "arm"), ("prop_crossproducts", "prop_crossproducts"),
                      ("Jinv", "Jinv")):
        np.testing.assert_allclose(getattr(m, attr), ma["%s_%s" % (name, key)], rtol=1e-13, atol=1e-18, err_msg=key)
    om = qo.model_constants(params)
    np.testing.assert_allclose(m.Jinv, om["Jinv"], rtol=1e-13)
    np.testing.assert_allclose(m.ou_sigma, om["ou_sigma"])


def test_low_inertia_params_match_reference():
    ma = load_npz("model_answers")
    assert _plain(quad_models.crazyflie_lowinertia_params()) == _plain(json.loads(str(ma["crazyflie_lowinertia_params_json"])))


def test_quadlink_full_inertia_matches_oracle():
    for p in (quad_models.crazyflie_params(), quad_models.defaultquad_params(), quad_models.mediumquad_params()):
        ql = QuadLink(p["geom"])
        oq = qo.quadlink_constants(p["geom"])
        np.testing.assert_allclose(ql.I_com, oq["I_com"], rtol=1e-13, atol=1e-20)
        np.testing.assert_allclose(ql.m, oq["mass"], rtol=1e-15)


def test_config_derived_values():
    m = ModelConstants(resolve_dynamics_params("DefaultQuad"))
    cfg, rc, sn = make_config(16, m, ep_time=5, sim_freq=200., sim_steps=2)
    assert cfg.ep_len == 500 and abs(cfg.dt - 0.005) < 1e-18                # 501-step episodes
    assert rc == {"pos": 1., "effort": 0.05, "action_change": 0., "crash": 1., "orient": 1., "yaw": 0., "rot": 0.,
                  "attitude": 0., "spin": 0.1, "vel": 0.}
    assert sn.bypass and cfg.sense.enabled == 0
    cfg, rc, sn = make_config(16, m, sense_noise="default", multi_reward=True, ep_time=4)
    assert cfg.sense.enabled == 1 and cfg.sense.pos_norm_std == 0.005 and cfg.sense.gyro_noise_density == 0.000175
    assert cfg.reward.log_pos == 1 and cfg.reward.effort == 0.01 and cfg.reward.spin == 0.0 and cfg.reward.pos_offset == 0.1
    assert cfg.model.ou_sigma == pytest.approx(0.01) and cfg.model.omega_max == 40.0
    assert [cfg.room_box[0][k] for k in range(3)] == [-10., -10., 0.] and [cfg.room_box[1][k] for k in range(3)] == [10.] * 3

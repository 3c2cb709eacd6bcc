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


# ---- dynamics randomisation: same np.random draws as the reference -------------------
def _random_cases():
    import copy
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    cases = []
    for seed in range(6):
        np.random.seed(100 + seed)
        cases.append(("randomquad_%d" % seed, qr.RandomQuad().sample()))
    base = quad_models.crazyflie_params()
    for seed, (ratio, sampler) in enumerate(((0.2, "normal"), (0.3, "uniform"), (0.1, "normal"))):
        np.random.seed(200 + seed)
        rs = qr.RelativeSampler(params=copy.deepcopy(base), noise_ratio=ratio, sampler=sampler,
                                noise_ratio_custom={"motor": {"linearity": 0.0}} if seed == 2 else None)
        cases.append(("relative_%d" % seed, rs.sample(copy.deepcopy(base))))
    np.random.seed(300)
    p = qr.sample_t2w_t2t(qr.sample_linearity(qr.sample_nodelay(quad_models.defaultquad_params())), 1.8, 2.5)
    cases.append(("helpers_0", p))
    cases.append(("simplified_cf", quad_models.crazyflie_params()))
    cases.append(("simplified_hb", quad_models.defaultquad_params()))
    np.random.seed(400)
    cases.append(("simplified_random", qr.sample_simplified_random_dyn()))
    return cases


def test_randomisation_reproduces_reference_draw_for_draw():
    """RandomQuad / RelativeSampler / sample_* under fixed seeds give the parameter dicts AND derived constants
    recorded from the live reference (tests/golden/random_models.npz, oracle/gen_golden.py random_models)."""
    rm = load_npz("random_models")
    cases = _random_cases()
    assert [n for n, _ in cases] == [str(n) for n in rm["names"]]
    for name, params in cases:
        ref_params = json.loads(str(rm[name + "_params_json"]))
        simp = name.startswith("simplified")                 # QuadLinkSimplified (inertia.py:312), dynamics_simplification=True
        m = ModelConstants(params, dynamics_simplification=simp)   # like the reference's links, fills in derived geometry
        _assert_tree_close(_plain(params), _plain(ref_params), name)
        for key in ("mass", "inertia", "prop_pos", "thrust_max", "torque_max", "arm", "Jinv"):
            np.testing.assert_allclose(getattr(m, key), rm["%s_%s" % (name, key)], rtol=1e-12, atol=1e-18,
                                       err_msg="%s %s" % (name, key))
        if simp:
            om = qo.model_constants(json.loads(str(rm[name + "_params_json"])), simplified=True)
            np.testing.assert_allclose(om["mass"], m.mass, rtol=1e-13)
            np.testing.assert_allclose(om["inertia"], m.inertia, rtol=1e-12)
            continue
        om = qo.model_constants(params)                      # the oracle's independent restatement agrees too
        np.testing.assert_allclose(om["mass"], m.mass, rtol=1e-13)
        np.testing.assert_allclose(om["inertia"], m.inertia, rtol=1e-12)
        np.testing.assert_allclose(om["Jinv"], m.Jinv, rtol=1e-10)


def _assert_tree_close(a, b, path):
    if isinstance(a, dict):
        assert set(a) == set(b), path
        for k in a:
            _assert_tree_close(a[k], b[k], path + "." + k)
    elif isinstance(a, list):
        np.testing.assert_allclose(np.array(a, dtype=float), np.array(b, dtype=float), rtol=1e-13, atol=0, err_msg=path)
    elif isinstance(a, float):
        assert a == pytest.approx(b, rel=1e-13, abs=0), path
    else:
        assert a == b, path


def test_randomquad_respects_limits_and_is_flyable():
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    np.random.seed(7)
    for _ in range(50):
        p = qr.RandomQuad().sample()
        m = ModelConstants(p)
        assert 1.2 <= p["motor"]["thrust_to_weight"] <= 3.5 and np.all(m.inertia > 0) and m.mass > 0
        assert np.all(p["motor"]["assymetry"] >= 0.9) and np.all(p["motor"]["assymetry"] <= 1.1)
        assert np.isfinite(m.Jinv).all() and abs(np.sum(m.thrust_max) / (9.81 * m.mass) - p["motor"]["thrust_to_weight"]) < 1e-9


def test_absolute_sampler_ranges():
    import copy
    import types
    from gym_art_b200.quadrotor import quadrotor_randomization as qr
    base = quad_models.crazyflie_params()
    rng = lambda lo, hi: types.SimpleNamespace(min=lo, max=hi)
    noise = copy.deepcopy(base)
    qr._walk(noise, lambda k, v: rng(np.min(v) * 0.9, np.max(v) * 1.1 + 1e-9) if not isinstance(v, str) else v)
    np.random.seed(1)
    p = qr.AbsoluteSampler(base, noise, sampler="uniform").sample(copy.deepcopy(base))
    assert 0.9 * 1.9 - 1e-9 <= p["motor"]["thrust_to_weight"] <= 1.1 * 1.9 + 1e-9
    assert p["motor"]["thrust_to_weight"] != base["motor"]["thrust_to_weight"]


def test_batched_model_pipeline_equals_the_scalar_one():
    """K random quadrotors through the vectorised sampler -> link model -> derived constants: every column of the
    [14, K, 4] table equals what the scalar pipeline (the one pinned on the reference) makes of the same parameters."""
    import time
    from gym_art_b200.quadrotor import batched_models as bm
    rng = np.random.RandomState(3)
    K, dt = 64, 0.005
    b = bm.sample_randomquad_batch(K, rng)
    rows = bm.model_rows_batch(b, dt)
    assert rows.shape == (14, K, 4)
    for k in range(K):
        want = ModelConstants(bm.params_of(b, k)).rows(dt)
        np.testing.assert_allclose(rows[:, k, :], want, rtol=1e-10, atol=1e-18, err_msg="model %d" % k)
    # same laws as randomquad_parameters (moments), and fast enough for one model per env
    big = bm.sample_randomquad_batch(20000, np.random.RandomState(4))
    t2w = big["thrust_to_weight"]
    assert abs(t2w.mean() - 2.5) < 0.02 and 1.5 <= t2w.min() and t2w.max() <= 3.5
    assert abs(big["geom"]["motor_xyz"][:, 0].mean() - 0.0625) < 0.001
    t0 = time.perf_counter()
    r = bm.model_rows_batch(big, dt)
    assert time.perf_counter() - t0 < 5.0 and np.isfinite(r).all() and (r[5, :, 3] > 0).all()

"""TEST INFRASTRUCTURE ONLY - generates tests/golden/*.npz from the LIVE reference.

Run in the build container (where /root/reference is mounted):

    python -m oracle.gen_golden            # rewrites every fixture

Each fixture is one single-env rollout of the unmodified reference QuadrotorEnv under a recorded
noise tape (oracle/ref_shim.NoiseTape): configuration, the unit noise draws regrouped into the
injected-noise slot layout of include/quadsim.h, the actions, and everything the reference
returned or held after every step.  The fixtures travel to the GPU box; the reference does not.
"""
import copy
import json
import os
import sys

import numpy as np

from . import quadsim_oracle as qo
from . import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

_uni = lambda lo, hi, n: (lambda rng: rng.uniform(lo, hi, n))


def _act_omega_thrust(rng):
    return np.array([rng.uniform(-0.2, 0.2), *rng.uniform(-1.0, 1.0, 3)])


def _act_vel_yaw(rng):
    return np.array([*rng.uniform(-1.0, 1.0, 3), rng.uniform(-0.5, 0.5)])


FULL_NOISE = dict(pos_unif_range=0.01, vel_unif_range=0.02, quat_norm_std=0.01, quat_unif_range=0.02, gyro_norm_std=0.1)
ALL_REW = dict(pos=1.0, effort=0.05, action_change=0.3, crash=1.0, orient=1.0, yaw=0.5, rot=0.2, attitude=0.3,
               spin=0.1, vel=0.2)
DRAG_CHANGE = {"motor": {"C_drag": 0.0028, "C_roll": 0.003, "linearity": 0.424, "damp_time_up": 0.1,
                         "damp_time_down": 0.3},
               "damp": {"vel": 0.001, "omega_quadratic": 0.015}}

# name -> (controller id, external controller name, action fn, steps, seed, multi, env kwargs)
CASES = {
    "hb_raw_nonoise": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 1100, 0, False,
                       dict(dynamics_params="DefaultQuad", raw_control=True, ep_time=5)),
    "cf_raw_noise_rnd": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 700, 1, False,
                         dict(dynamics_params="Crazyflie", raw_control=True, ep_time=3, sense_noise="default",
                              init_random_state=True)),
    "cf_omegathrust_noise_rnd": (qo.CTRL_OMEGA_THRUST, "omega_thrust", _act_omega_thrust, 700, 2, False,
                                 dict(dynamics_params="Crazyflie", ep_time=3, sense_noise="default",
                                      init_random_state=True)),
    "hb_mellinger_noise_act": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 1100, 3, False,
                               dict(dynamics_params="DefaultQuad", raw_control=False, ep_time=5, sense_noise="default",
                                    obs_repr="xyz_vxyz_R_omega_act")),
    "hb_mellinger_nonoise": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 1100, 4, False,
                             dict(dynamics_params="DefaultQuad", raw_control=False, ep_time=5)),
    "cf_mellinger_rnd": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 500, 5, False,
                         dict(dynamics_params="Crazyflie", raw_control=False, ep_time=2, init_random_state=True)),
    "cf_fullnoise_accact": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 6, False,
                            dict(dynamics_params="Crazyflie", ep_time=1, sense_noise=FULL_NOISE,
                                 obs_repr="xyz_vxyz_R_omega_acc_act")),
    "hb_velyaw": (qo.CTRL_VELOCITY_YAW, "velocity_yaw", _act_vel_yaw, 700, 7, False,
                  dict(dynamics_params="DefaultQuad", ep_time=3)),
    "hb_shifted": (qo.CTRL_SHIFTED, "shifted", _uni(-1, 1.8, 4), 300, 8, False,
                   dict(dynamics_params="DefaultQuad", ep_time=1)),
    "hb_vertical": (qo.CTRL_VERTICAL, None, _uni(-1, 1, 1), 300, 9, False,
                    dict(dynamics_params="DefaultQuad", dim_mode="1D", ep_time=1)),
    "hb_raw_nozm": (qo.CTRL_RAW, None, _uni(0, 1, 4), 300, 10, False,
                    dict(dynamics_params="DefaultQuad", raw_control_zero_middle=False, ep_time=1)),
    "hb_raw_allrew": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 11, False,
                      dict(dynamics_params="DefaultQuad", ep_time=1, rew_coeff=ALL_REW)),
    "hb_multi_reward": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 12, True,
                        dict(dynamics_params="DefaultQuad", ep_time=1)),
    "cf_drag_damp": (qo.CTRL_RAW, None, _uni(-0.2, 0.6, 4), 300, 13, False,
                     dict(dynamics_params="Crazyflie", ep_time=1, dynamics_change=DRAG_CHANGE)),
    "hb_xyzr_h_noise": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 14, False,
                        dict(dynamics_params="DefaultQuad", ep_time=1, sense_noise="default",
                             obs_repr="xyzr_vxyzr_R_omega_h")),
    # dynamics randomisation: the reference samples a RandomQuad / perturbs the CrazyFlie itself
    "rq_raw_noise": (qo.CTRL_RAW, None, _uni(-0.6, 1.0, 4), 300, 16, False,
                     dict(dynamics_params="RandomQuad", ep_time=1, sense_noise="default")),
    "rq_mellinger": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 450, 17, False,
                     dict(dynamics_params="RandomQuad", raw_control=False, ep_time=2, init_random_state=True)),
    "cf_relative_omegathrust": (qo.CTRL_OMEGA_THRUST, "omega_thrust", _act_omega_thrust, 300, 18, False,
                                dict(dynamics_params="Crazyflie", ep_time=1,
                                     dyn_sampler_1={"class": "RelativeSampler", "noise_ratio": 0.2, "sampler": "normal"})),
    # moving goals: excite (a new random goal every 5 ticks, quadrotor.py:957-962) and resample_goal (:1078-1079)
    "hb_mellinger_excite": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 450, 19, False,
                            dict(dynamics_params="DefaultQuad", raw_control=False, ep_time=2, excite=True, sense_noise="default")),
    "cf_raw_resample_goal": (qo.CTRL_RAW, None, _uni(-0.2, 0.6, 4), 300, 20, False,
                             dict(dynamics_params="Crazyflie", ep_time=0.5, resample_goal=True)),
    "cf_raw_simplified": (qo.CTRL_RAW, None, _uni(-0.2, 0.8, 4), 300, 21, False,
                          dict(dynamics_params="Crazyflie", ep_time=1, dynamics_simplification=True, sense_noise="default")),
    # observation builders that run once get_state.py is given the names it forgot to import (oracle/ref_shim.py)
    "hb_quat_fullnoise": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 22, False,
                          dict(dynamics_params="DefaultQuad", ep_time=1, sense_noise=FULL_NOISE, obs_repr="xyz_vxyz_quat_omega")),
    "cf_xyzr_quat_h_rnd": (qo.CTRL_RAW, None, _uni(-0.2, 0.8, 4), 300, 23, False,
                           dict(dynamics_params="Crazyflie", ep_time=1, sense_noise="default", init_random_state=True,
                                obs_repr="xyzr_vxyzr_quat_omega_h")),
    "hb_t2w_t2t_nonoise": (qo.CTRL_RAW, None, _uni(-1, 1, 4), 300, 24, False,
                           dict(dynamics_params="DefaultQuad", ep_time=1, sense_noise=None, obs_repr="xyz_vxyz_R_omega_t2w_t2t")),
    "hb_xyzr_t2w_mellinger": (qo.CTRL_MELLINGER, None, _uni(0, 0, 4), 450, 25, False,
                              dict(dynamics_params="DefaultQuad", raw_control=False, ep_time=2, sense_noise="default",
                                   obs_repr="xyzr_vxyzr_R_omega_t2w", t2w_std=0.4)),
    # init_random_state outside 3D (quadrotor.py:1101-1103)
    "hb_vertical_rnd": (qo.CTRL_VERTICAL, None, _uni(-1, 1, 1), 300, 26, False,
                        dict(dynamics_params="DefaultQuad", dim_mode="1D", ep_time=0.5, init_random_state=True)),
    "mq_raw_simsteps4": (qo.CTRL_RAW, None, _uni(-0.5, 0.8, 4), 300, 15, False,
                         dict(dynamics_params="MediumQuad", ep_time=2, sim_steps=4, sim_freq=100.0)),
}


def _plain(x):
    if isinstance(x, dict):
        return {k: _plain(v) for k, v in x.items()}
    if isinstance(x, np.ndarray):
        return x.tolist()
    if isinstance(x, (np.floating, np.integer)):
        return x.item()
    if isinstance(x, (list, tuple)):
        return [_plain(v) for v in x]
    return x


def rollout(name):
    ctrl_id, ext, act_fn, n_steps, seed, multi, kw = CASES[name]
    rr = rh.ReferenceRunner(seed=seed, controller=ext, multi=multi, **kw)
    env = rr.env
    A = qo.CTRL_ACT_DIM[ctrl_id]
    rng = np.random.RandomState(1000 + seed)
    init = dict(ou0=env.dynamics.thrust_noise.state.copy(), svd_timer0=float(env.dynamics.since_last_svd),
                gyro_bias0=env.sense_noise.gyro_bias.copy())
    rec = {k: [] for k in ("actions", "ou_z", "obs_z", "obs", "rew", "done", "pos", "vel", "rot", "omega", "rot_damp",
                           "cmds_damp", "ou", "accelerometer", "reset_z", "reset_obs", "reset_step", "goal_z", "reset_goal_z",
                           "goal")}
    obs, rz = rr.reset()
    rec["reset_goal_z"].append(rr.last_goal_z.copy())
    rec["reset_z"].append(rz)
    rec["reset_obs"].append(obs)
    rec["reset_step"].append(0)
    for t in range(n_steps):
        a = act_fn(rng)
        a4 = np.zeros(4)
        a4[:A] = a
        o, r, d, info, ou_z, obs_z = rr.step(a4)
        st = rr.state()
        for k, v in (("actions", a), ("ou_z", ou_z), ("obs_z", obs_z), ("obs", o), ("rew", r), ("done", d),
                     ("goal_z", rr.last_goal_z.copy()), ("goal", np.array(env.goal, dtype=np.float64))):
            rec[k].append(v)
        for k in ("pos", "vel", "rot", "omega", "rot_damp", "cmds_damp", "ou", "accelerometer"):
            rec[k].append(st[k])
        if d:
            obs, rz = rr.reset()
            rec["reset_goal_z"].append(rr.last_goal_z.copy())
            rec["reset_z"].append(rz)
            rec["reset_obs"].append(obs)
            rec["reset_step"].append(t + 1)
    out = {k: np.array(v) for k, v in rec.items()}
    out.update(init)
    meta = dict(name=name, controller=ctrl_id, multi=multi, env_kwargs=_plain(kw), params=_plain(env.dynamics_params),
                oracle_kwargs=_plain(rh.oracle_kwargs_from_env(env, ctrl_id)), ep_len=env.ep_len,
                numpy=np.__version__, reference="amolchanov86/gym_art (live, under oracle/ref_shim.py stubs)")
    out["meta_json"] = np.array(json.dumps(meta))
    return out


def known_answer():
    """SURVEY.md section 8c deterministic known-answer case, regenerated rather than transcribed."""
    out = {}
    for tag, quad in (("hb", "DefaultQuad"), ("cf", "Crazyflie")):
        rr = rh.ReferenceRunner(seed=0, dynamics_params=quad, raw_control=True, ep_time=5,
                                dynamics_change={"noise": {"thrust_noise_ratio": 0.0}})
        env = rr.env
        yaw = 0.3
        R = np.array([[np.cos(yaw), -np.sin(yaw), 0.0], [np.sin(yaw), np.cos(yaw), 0.0], [0.0, 0.0, 1.0]])
        env.dynamics.set_state(np.array([0.5, -0.25, 1.5]), np.array([0.1, 0.2, -0.3]), R, np.array([0.5, -0.4, 0.3]))
        env.dynamics.reset()
        env.dynamics.since_last_svd = 0
        env.tick = 0
        env.actions = [np.zeros(4), np.zeros(4)]
        a = np.array([0.1, -0.2, 0.3, 0.0])
        for t in range(100):
            o, r, d, info, ou_z, obs_z = rr.step(a)
            if t in (0, 99):
                st = rr.state()
                for k in ("pos", "vel", "rot", "omega", "cmds_damp"):
                    out["%s_%d_%s" % (tag, t + 1, k)] = st[k]
                out["%s_%d_rew" % (tag, t + 1)] = np.array(r)
                out["%s_%d_obs" % (tag, t + 1)] = o
                out["%s_%d_crashed" % (tag, t + 1)] = np.array(bool(env.crashed))
        out["%s_params_json" % tag] = np.array(json.dumps(_plain(env.dynamics_params)))
    out["R0"] = R
    return out


def model_answers():
    """Host-constant known answers for every shipped parameter set."""
    ref = rh.ref_shim.import_reference()
    out = {}
    for name in ("DefaultQuad", "Crazyflie", "MediumQuad"):
        rr = rh.ReferenceRunner(seed=0, dynamics_params=name, raw_control=False)
        d = rr.env.dynamics
        for k in ("mass", "inertia", "prop_pos", "thrust_max", "torque_max", "arm", "prop_crossproducts"):
            out["%s_%s" % (name, k)] = np.array(getattr(d, k), dtype=np.float64)
        out["%s_Jinv" % name] = rr.env.controller.Jinv
        out["%s_params_json" % name] = np.array(json.dumps(_plain(rr.env.dynamics_params)))
    low = ref.quad_models.crazyflie_lowinertia_params()
    out["crazyflie_lowinertia_params_json"] = np.array(json.dumps(_plain(low)))
    return out


def random_models():
    """Dynamics randomisation: parameter dicts and derived constants the live
    reference produces under fixed np.random seeds - RandomQuad samples, RelativeSampler perturbations of the
    CrazyFlie (normal and uniform), and the sample_* helpers."""
    ref = rh.ref_shim.import_reference()
    qr = ref.quad_rand
    out, cases = {}, []
    for seed in range(6):
        np.random.seed(100 + seed)
        cases.append(("randomquad_%d" % seed, qr.RandomQuad().sample()))
    base = ref.quad_models.crazyflie_params()
    for seed, (ratio, sampler) in enumerate(((0.2, "normal"), (0.3, "uniform"), (0.1, "normal"))):
        np.random.seed(200 + seed)
        rs = qr.RelativeSampler(params=copy.deepcopy(base), noise_ratio=ratio, sampler=sampler,
                                noise_ratio_custom={"motor": {"linearity": 0.0}} if seed == 2 else None)
        cases.append(("relative_%d" % seed, rs.sample(copy.deepcopy(base))))
    np.random.seed(300)
    p = qr.sample_t2w_t2t(qr.sample_linearity(qr.sample_nodelay(ref.quad_models.defaultquad_params())), 1.8, 2.5)
    cases.append(("helpers_0", p))
    for simp_name, simp_params in (("simplified_cf", ref.quad_models.crazyflie_params()),
                                   ("simplified_hb", ref.quad_models.defaultquad_params())):
        cases.append((simp_name, simp_params))
    np.random.seed(400)
    cases.append(("simplified_random", qr.sample_simplified_random_dyn()))
    for name, params in cases:
        d = ref.quadrotor.QuadrotorDynamics(model_params=params, dynamics_steps_num=2,
                                            dynamics_simplification=name.startswith("simplified"))
        ctrl = ref.control.OmegaThrustControl(d)
        out[name + "_params_json"] = np.array(json.dumps(_plain(params)))
        for k in ("mass", "inertia", "prop_pos", "thrust_max", "torque_max", "arm"):
            out["%s_%s" % (name, k)] = np.array(getattr(d, k), dtype=np.float64)
        out[name + "_Jinv"] = ctrl.Jinv
    out["names"] = np.array([n for n, _ in cases])
    return out


def affine_answers():
    """F(s, dt) and G(s) of the live reference (quadrotor.py:461-518) on random states: both shipped models with the damping
    terms switched on, all three dim_modes."""
    import contextlib
    import io
    ref = rh.ref_shim.import_reference()
    rng = np.random.RandomState(5)
    out = {}
    for tag, name in (("hb", "defaultquad_params"), ("cf", "crazyflie_params")):
        params = getattr(ref.quad_models, name)()
        params["damp"]["omega_quadratic"], params["damp"]["vel"] = 0.015, 0.001
        out[tag + "_params_json"] = np.array(json.dumps(_plain(params)))
        S = rng.uniform(-3, 3, (16, 21))
        S[:, 15:18] *= 4.0
        out[tag + "_s"] = S
        for dm in ("1D", "2D", "3D"):
            with contextlib.redirect_stdout(io.StringIO()):
                d = ref.quadrotor.QuadrotorDynamics(model_params=params, dynamics_steps_num=2, dim_mode=dm)
            out["%s_%s_F" % (tag, dm)] = np.stack([d.F(s, 0.005) for s in S])
            out["%s_%s_G" % (tag, dm)] = np.stack([d.G(s) for s in S])
    out["dt"] = np.array(0.005)
    return out


def restricted_state_answers():
    """QuadrotorDynamics.pitch_roll_restricted_random_state (quadrotor.py:243-259) under a NoiseTape: the unit draws in
    the reset_z layout (pitch, roll, yaw in the RZ_NUP slots) and what the reference returned."""
    import contextlib
    import io
    ref = rh.ref_shim.import_reference()
    params = ref.quad_models.crazyflie_params()
    with contextlib.redirect_stdout(io.StringIO()):
        d = ref.quadrotor.QuadrotorDynamics(model_params=params, dynamics_steps_num=2)
    tape = rh.ref_shim.NoiseTape(ref, seed=65)
    kw = dict(pitch_max=0.4, roll_max=0.7, yaw_max=2.5)
    n = 64
    rz, vel, rot, omega = np.zeros((n, qo.RESET_Z)), np.zeros((n, 3)), np.zeros((n, 3, 3)), np.zeros((n, 3))
    for i in range(n):
        with tape:
            _, vel[i], rot[i], omega[i] = d.pitch_roll_restricted_random_state(box=2.0, vel_max=3.0, omega_max=5.0, **kw)
        log = [v for k, v in tape.take()]
        rz[i, qo.RZ_UPOS:qo.RZ_UPOS + 3] = log[0]
        rz[i, qo.RZ_UVEL:qo.RZ_UVEL + 3], rz[i, qo.RZ_UVELMAG] = log[1], log[2]
        rz[i, qo.RZ_UOM:qo.RZ_UOM + 3], rz[i, qo.RZ_UOMMAG] = log[3], log[4]
        rz[i, qo.RZ_NUP:qo.RZ_NUP + 3] = log[5], log[6], log[7]
    return dict(reset_z=rz, vel=vel, rot=rot, omega=omega, params_json=np.array(json.dumps(_plain(params))),
                vel_max=np.array(3.0), omega_max=np.array(5.0), **{k: np.array(v) for k, v in kw.items()})


ANSWER_FILES = {"known_answer": known_answer, "model_answers": model_answers, "random_models": random_models,
                "affine_answers": affine_answers, "restricted_state": restricted_state_answers}


def main(argv):
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    names = argv[1:] or list(CASES)
    for name in names:
        out = rollout(name)
        np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **out)
        print("wrote", name, {k: v.shape for k, v in out.items() if k in ("obs", "reset_z")})
    if not argv[1:]:
        for name, fn in ANSWER_FILES.items():
            np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **fn())
        print("wrote", ", ".join(ANSWER_FILES))


if __name__ == "__main__":
    main(sys.argv)

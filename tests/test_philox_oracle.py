"""CPU: the NumPy Philox restatement (oracle/philox.py) - Random123 known-answer vectors, the uniform / normal
transforms, and the injected-noise tape it lays out for OracleVecEnv."""
import numpy as np

from oracle import philox as ph
from oracle import quadsim_oracle as qo

# Random123 kat_vectors, philox4x32 with 10 rounds: (counter, key) -> output
KAT = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
       ((0xffffffff,) * 4, (0xffffffff, 0xffffffff), (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
       ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0), (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]


def test_philox4x32_10_known_answers():
    for ctr, key, want in KAT:
        got = ph.philox4x32_10(np.array(ctr, dtype=np.uint32), key)
        assert tuple(int(x) for x in got) == want
    batch = ph.philox4x32_10(np.array([k[0] for k in KAT[:1] * 5], dtype=np.uint32), KAT[0][1])
    assert batch.shape == (5, 4) and (batch == np.array(KAT[0][2], dtype=np.uint32)).all()


def test_counter_layout_separates_env_step_and_slot():
    ids = np.array([0, 1, (1 << 32) + 1, (1 << 40) + 7], dtype=np.uint64)
    a = ph.raw(5, ids, 3, 64)
    assert len({tuple(r) for r in a.tolist()}) == 4                      # the high word of the env id is part of the counter
    assert not (ph.raw(5, ids, 4, 64) == a).all(axis=1).any()             # step
    assert not (ph.raw(5, ids, 3, 65) == a).all(axis=1).any()             # slot
    assert not (ph.raw(6, ids, 3, 64) == a).all(axis=1).any()             # seed (key)
    assert not (ph.raw(5 + (1 << 32), ids, 3, 64) == a).all(axis=1).any()  # high half of the seed
    assert not (ph.raw(5, ids, 3 + (1 << 32), 64) == a).all(axis=1).any()  # high bits of the step


def test_transforms_have_the_right_laws():
    ids = np.arange(200000, dtype=np.uint64)
    w = ph.raw(1, ids, 0, 64)
    u = ph.unif4(w)
    assert u.min() >= -1.0 and u.max() < 1.0 and abs(u.mean()) < 5e-3 and abs(u.var() - 1 / 3) < 5e-3
    for fp32 in (True, False):
        z = ph.normal4(w, fp32)
        assert np.isfinite(z).all() and abs(z.mean()) < 5e-3 and abs(z.var() - 1.0) < 1e-2
        assert abs(np.mean(z ** 4) - 3.0) < 0.1
        assert abs(np.corrcoef(z[:, 0], z[:, 1])[0, 1]) < 1e-2
    assert np.median(np.abs(ph.normal4(w, True) - ph.normal4(w, False))) < 1e-6   # 24- vs 32-bit mantissas of the same words


def test_tape_layout():
    ids = np.arange(1000, 1300, dtype=np.uint64)
    tape = ph.PhiloxTape(seed=9, fp32=False, sim_steps=3, init_random_state=True)
    assert tape.ou_z(ids, 4).shape == (300, 3, 4) and tape.obs_z(ids, 4).shape == (300, qo.OBS_Z)
    rz = tape.reset_z(ids, 4)
    assert rz.shape == (300, qo.RESET_Z)
    # obs group g = first three words of slot g: normal or uniform as the layout says
    z = tape.obs_z(ids, 4)
    assert np.array_equal(z[:, qo.Z_UPOS:qo.Z_UPOS + 3], ph.unif4(ph.raw(9, ids, 4, 1))[:, :3])
    assert np.array_equal(z[:, qo.Z_NACC_D:qo.Z_NACC_D + 3], ph.normal4(ph.raw(9, ids, 4, 9), False)[:, :3])
    assert np.array_equal(rz[:, qo.RZ_OBS + qo.Z_NVEL:qo.RZ_OBS + qo.Z_NVEL + 3], ph.normal4(ph.raw(9, ids, 4, 18), False)[:, :3])
    # default-spawn yaw: within 60 degrees of the direction from the spawn point to the origin (quadrotor.py:1124-1126)
    t2 = ph.PhiloxTape(seed=9, fp32=False)
    rz = t2.reset_z(ids, 0)
    x, y = 2.0 * rz[:, 0], 2.0 * rz[:, 1]
    psi = np.pi * rz[:, qo.RZ_YAW]
    head = np.stack([np.cos(psi), np.sin(psi)], axis=1)
    to_origin = -np.stack([x, y], axis=1) / np.linalg.norm(np.stack([x, y], axis=1), axis=1, keepdims=True)
    assert np.all(np.sum(head * to_origin, axis=1) >= 0.5 - 1e-12)
    # rand_uniform_rot3d: the accepted forward vector is never within acos(0.95) of `up`
    t3 = ph.PhiloxTape(seed=2, fp32=False, init_random_state=True)
    rz = t3.reset_z(np.arange(50000, dtype=np.uint64), 1)
    up, fw = rz[:, qo.RZ_NUP:qo.RZ_NUP + 3], rz[:, qo.RZ_NFWD:qo.RZ_NFWD + 3]
    cosang = np.sum(up * fw, axis=1) / np.linalg.norm(up, axis=1) / np.linalg.norm(fw, axis=1)
    assert cosang.max() <= 0.95

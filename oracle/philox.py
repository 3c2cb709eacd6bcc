"""TEST INFRASTRUCTURE ONLY - NumPy restatement of the device's Philox noise source (csrc/qs_device.cuh:264-307, slot map
:281-283, transforms Mth<R>::normal2 / unif :184-214), so that the PHILOX-mode kernels - the ones bench.py times - can be
compared with OracleVecEnv draw for draw: the oracle is fed, through its injected-noise arguments, exactly the unit draws
the kernel generates for (seed, global env id, step index, slot).

Philox4x32-10 is Salmon, Moraes, Dror, Shaw, "Parallel random numbers: as easy as 1, 2, 3" (SC'11); the known-answer
vectors of the Random123 distribution (kat_vectors) pin `philox4x32_10` below in tests/test_philox_oracle.py.

    key     = (seed & 0xffffffff, seed >> 32)
    counter = (gid & 0xffffffff, gid >> 32, step & 0xffffffff, ((step >> 32) & 0xffffff) | slot << 24)

Integer words and the uniform slots are bit-exact.  Normals: the fp64 kernels use the same libm-grade formula as here; the
fp32 throughput kernels use MUFU approximations (lg2.approx, rsqrt.approx, sin/cos.approx) of it - a few fp32 ulp.
"""
import numpy as np

from . import quadsim_oracle as qo

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)

# slot map (qs_device.cuh:281-283)
SLOT_OBS, SLOT_EXCITE, SLOT_RESET_OBS, SLOT_RESET_POS, SLOT_RESET_VEL, SLOT_RESET_OM = 0, 12, 16, 26, 27, 28
SLOT_RESET_UP, SLOT_RESET_FWD, SLOT_RESET_GOAL, SLOT_OU = 29, 30, 40, 64
SLOT_T2W, SLOT_RESET_T2W = 11, 41      # group 10 of an observation call: t2w / t2t normals (QS_Z_NT2W, QS_Z_NT2T)
# groups of one add_noise call, in OBS_Z order, and whether the group is uniform (quadsim.h QS_Z_*)
OBS_GROUP_UNIFORM = (False, True, False, True, False, False, False, True, False, False)


def philox4x32_10(counter, key):
    """counter: uint32 [..., 4]; key: (k0, k1) python ints.  Returns uint32 [..., 4]."""
    c = np.asarray(counter, dtype=np.uint64)
    c0, c1, c2, c3 = c[..., 0], c[..., 1], c[..., 2], c[..., 3]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    for _ in range(10):
        p0 = M0 * c0                                  # 64-bit products of 32-bit operands: no overflow in uint64
        p1 = M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & MASK
        hi1, lo1 = p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0 = (k0 + W0) & 0xFFFFFFFF
        k1 = (k1 + W1) & 0xFFFFFFFF
    return np.stack([c0, c1, c2, c3], axis=-1).astype(np.uint32)


def raw(seed, gids, step, slot):
    """The four 32-bit words the kernel draws for every global env id in `gids` at (step, slot)."""
    gids = np.asarray(gids, dtype=np.uint64)
    step = int(step)
    c = np.empty(gids.shape + (4,), dtype=np.uint64)
    c[..., 0] = gids & MASK
    c[..., 1] = gids >> np.uint64(32)
    c[..., 2] = step & 0xFFFFFFFF
    c[..., 3] = ((step >> 32) & 0x00FFFFFF) | (int(slot) << 24)
    return philox4x32_10(c, (int(seed) & 0xFFFFFFFF, (int(seed) >> 32) & 0xFFFFFFFF))


def unif4(words):
    """Mth<R>::unif: int32(word) * 2^-31, in [-1, 1)."""
    return words.astype(np.int32).astype(np.float64) * 4.656612873077393e-10


def normal4(words, fp32):
    """Mth<R>::normal2 on (x, y) and (z, w): Box-Muller pairs (z0, z1) = r (cos, sin)."""
    w = words.astype(np.uint64)
    out = np.empty(words.shape, dtype=np.float64)
    for a, b, k in ((w[..., 0], w[..., 1], 0), (w[..., 2], w[..., 3], 2)):
        if fp32:                                                          # 24-bit mantissas, qs_device.cuh:185-188
            u1 = ((a >> np.uint64(8)) + np.uint64(1)).astype(np.float64) * 2.0 ** -24
            ang = (b >> np.uint64(8)).astype(np.float64) * (2.0 * np.pi * 2.0 ** -24)
        else:                                                             # qs_device.cuh:206-210
            u1 = (a.astype(np.float64) + 1.0) * 2.0 ** -32
            ang = 2.0 * np.pi * (b.astype(np.float64) * 2.0 ** -32)
        r = np.sqrt(-2.0 * np.log(u1))
        out[..., k] = r * np.cos(ang)
        out[..., k + 1] = r * np.sin(ang)
    return out


class PhiloxTape(object):
    """Unit draws of a Philox-mode handle in the injected-noise layout OracleVecEnv consumes."""

    def __init__(self, seed, fp32=True, sim_steps=2, init_random_state=False, goal=(0.0, 0.0, 2.0), box=2.0, dim_mode="3D"):
        self.seed, self.fp32, self.sim_steps = int(seed), bool(fp32), int(sim_steps)
        # False / True / "pitch_roll_restricted" (QsInitMode)
        self.init_random_state = bool(init_random_state)
        self.restricted = init_random_state == "pitch_roll_restricted"
        self.goal, self.box, self.dim_mode = np.asarray(goal, dtype=np.float64), float(box), dim_mode

    def ou_z(self, gids, step):
        """[n, sim_steps, 4] normals of OUNoise.noise(), one Philox call per substep (slot 64 + substep)."""
        return np.stack([normal4(raw(self.seed, gids, step, SLOT_OU + s), self.fp32) for s in range(self.sim_steps)], axis=1)

    def obs_z(self, gids, step, reset_obs=False):
        """[n, OBS_Z]: group g of the add_noise call = first three words of slot (0 | 16) + g."""
        base = SLOT_RESET_OBS if reset_obs else SLOT_OBS
        z = np.zeros((len(gids), qo.OBS_Z))
        for g, is_unif in enumerate(OBS_GROUP_UNIFORM):
            w = raw(self.seed, gids, step, base + g)
            z[:, 3 * g:3 * g + 3] = (unif4(w) if is_unif else normal4(w, self.fp32))[:, :3]
        w = raw(self.seed, gids, step, SLOT_RESET_T2W if reset_obs else SLOT_T2W)
        z[:, 30:33] = normal4(w, self.fp32)[:, :3]
        return z

    def reset_z(self, gids, step, goals=None):
        """[n, RESET_Z] for a reset that happens at Philox step index `step` (an explicit qs_reset, or the reset folded
        into the env step with that index).  goals: per-env goal [n, 3] (default: the handle's goal)."""
        n = len(gids)
        z = np.zeros((n, qo.RESET_Z))
        u = unif4(raw(self.seed, gids, step, SLOT_RESET_POS))
        z[:, qo.RZ_UPOS:qo.RZ_UPOS + 3] = u[:, :3]
        g = np.tile(self.goal, (n, 1)) if goals is None else np.asarray(goals, dtype=np.float64).reshape(n, 3)
        if self.fp32:      # the kernel forms the spawn point in fp32 before taking atan2 of it (qs_device.cuh reset_env)
            x = (np.float32(self.box) * u[:, 0].astype(np.float32) + g[:, 0].astype(np.float32)).astype(np.float64)
            y = (np.float32(self.box) * u[:, 1].astype(np.float32) + g[:, 1].astype(np.float32)).astype(np.float64)
        else:
            x, y = self.box * u[:, 0] + g[:, 0], self.box * u[:, 1] + g[:, 1]
        if self.dim_mode == "1D":
            x, y = g[:, 0], g[:, 1]
        elif self.dim_mode == "2D":
            y = g[:, 1]
        # Philox mode draws the law of the reference's rejection loop directly: psi = atan2(-y, -x) + (pi / 3) u
        # (DESIGN.md section 5.2); the oracle's slot is "accepted yaw / pi"
        z[:, qo.RZ_YAW] = (np.arctan2(-y, -x) + (np.pi / 3.0) * u[:, 3]) / np.pi
        if self.init_random_state and self.dim_mode != "3D":              # quadrotor.py:1101-1110: unit draws standing for rand()
            z[:, qo.RZ_UVEL:qo.RZ_UVEL + 4] = unif4(raw(self.seed, gids, step, SLOT_RESET_VEL))
            z[:, qo.RZ_UOM:qo.RZ_UOM + 4] = unif4(raw(self.seed, gids, step, SLOT_RESET_OM))
            z[:, qo.RZ_YAW] = u[:, 3]                                     # 2D: the pitch angle draw
        elif self.restricted:                                             # quadrotor.py:243-259: pitch, roll, yaw uniforms
            z[:, qo.RZ_UVEL:qo.RZ_UVEL + 4] = unif4(raw(self.seed, gids, step, SLOT_RESET_VEL))
            z[:, qo.RZ_UOM:qo.RZ_UOM + 4] = unif4(raw(self.seed, gids, step, SLOT_RESET_OM))
            z[:, qo.RZ_NUP:qo.RZ_NUP + 3] = unif4(raw(self.seed, gids, step, SLOT_RESET_UP))[:, :3]
        elif self.init_random_state:
            a = unif4(raw(self.seed, gids, step, SLOT_RESET_VEL))
            b = unif4(raw(self.seed, gids, step, SLOT_RESET_OM))
            z[:, qo.RZ_UVEL:qo.RZ_UVEL + 4] = a
            z[:, qo.RZ_UOM:qo.RZ_UOM + 4] = b
            up = normal4(raw(self.seed, gids, step, SLOT_RESET_UP), self.fp32)[:, :3]
            z[:, qo.RZ_NUP:qo.RZ_NUP + 3] = up
            upn = up / np.maximum(np.linalg.norm(up, axis=1, keepdims=True), 1e-300)
            pending = np.ones(n, dtype=bool)
            for tr in range(8):                                           # rand_uniform_rot3d's rejection loop, bounded
                fw = normal4(raw(self.seed, gids, step, SLOT_RESET_FWD + tr), self.fp32)[:, :3]
                z[pending, qo.RZ_NFWD:qo.RZ_NFWD + 3] = fw[pending]
                fwn = fw / np.maximum(np.linalg.norm(fw, axis=1, keepdims=True), 1e-300)
                pending &= np.sum(fwn * upn, axis=1) > 0.95
                if not pending.any():
                    break
        z[:, qo.RZ_OBS:] = self.obs_z(gids, step, reset_obs=True)
        return z

    def goal_z(self, gids, step, at_reset):
        return unif4(raw(self.seed, gids, step, SLOT_RESET_GOAL if at_reset else SLOT_EXCITE))

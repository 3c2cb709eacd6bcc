"""GPU: libquadsim.so driven from a plain C program (tests/c/abi_smoke.c: dlopen + cudaMalloc, no Python, no torch)
gives the same numbers as the Python binding, and the C header and the ctypes mirror agree on sizeof(QsConfig)."""
import ctypes as C
import os
import shutil
import subprocess

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_c_program_matches_python_binding(tmp_path):
    import torch
    from gym_art_b200 import QuadVecEnv, _lib
    cuda = os.environ.get("CUDA_HOME", "/usr/local/cuda")
    exe = str(tmp_path / "abi_smoke")
    cc = shutil.which("gcc") or shutil.which("cc")
    r = subprocess.run([cc, os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-I" + os.path.join(ROOT, "include"),
                        "-I" + cuda + "/include", "-L" + cuda + "/lib64", "-lcudart", "-ldl", "-lm", "-o", exe],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    n, steps = 10000 + 7, 220
    env = QuadVecEnv(n, dynamics_params="DefaultQuad", controller="mellinger", ep_time=1, seed=5, sense_noise="default")
    (tmp_path / "config.bin").write_bytes(bytes(env.cfg))
    env.reset()
    for _ in range(steps):
        o, rew, d = env.step(None)
    torch.cuda.synchronize()
    want = dict(mean_dist=float(o[:, :3].double().norm(dim=1).mean()), mean_rew=float(rew.double().mean()),
                obs_sum=float(o.double().sum()), episodes=env.stats()[2])
    env_vars = dict(os.environ, LD_LIBRARY_PATH=cuda + "/lib64:" + os.environ.get("LD_LIBRARY_PATH", ""))
    r = subprocess.run([exe, _lib.LIB_PATH, str(tmp_path / "config.bin"), str(steps)], capture_output=True, text=True, env=env_vars)
    assert r.returncode == 0, r.stdout + r.stderr
    tok = r.stdout.split()
    got = {tok[k]: float(tok[k + 1]) for k in range(0, len(tok), 2)}
    assert got["n"] == n and got["obs_dim"] == 18 and got["rows"] == env.layout.rows
    assert got["episodes"] == want["episodes"] == 2 * n               # 101-step episodes
    for k in ("mean_dist", "mean_rew", "obs_sum"):                    # same kernels, same Philox streams: same numbers
        assert got[k] == pytest.approx(want[k], rel=1e-6, abs=1e-6), k
    assert C.sizeof(_lib.QsConfig) == os.path.getsize(tmp_path / "config.bin")

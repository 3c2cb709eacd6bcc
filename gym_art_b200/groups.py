"""Several independent env groups stepped concurrently on their own CUDA streams.

A single mid-size batch (tens of thousands of envs) is one wave of CTAs: its load, compute and store phases cannot
overlap each other, so a B200 steps it at about half of its HBM roofline.  When the work is
several such batches that do not depend on one another - parallel samplers, populations, hyper-parameter sweeps -
stepping them on separate streams lets one group's loads run under another group's compute and stores.  Each group is
an ordinary QuadVecEnv (own handle, own Philox stream ids); nothing here changes what a group computes.
"""


class QuadVecEnvGroup(object):
    def __init__(self, num_groups, num_envs, device="cuda:0", **env_kwargs):
        import torch
        from .vec_env import QuadVecEnv
        self.torch = torch
        self.device = torch.device(device)
        seed = env_kwargs.pop("seed", 0)
        base = env_kwargs.pop("env_id_offset", 0)
        self.envs = [QuadVecEnv(num_envs, device=device, seed=seed, env_id_offset=base + g * num_envs, **env_kwargs)
                     for g in range(num_groups)]
        self.streams = [torch.cuda.Stream(device=self.device) for _ in range(num_groups)]
        self.num_groups, self.num_envs = num_groups, num_envs

    def _fanout(self, fn):
        t = self.torch
        cur = t.cuda.current_stream(self.device)
        outs = []
        for g, (env, st) in enumerate(zip(self.envs, self.streams)):
            st.wait_stream(cur)
            with t.cuda.stream(st):
                outs.append(fn(g, env))
        for st in self.streams:
            cur.wait_stream(st)
        return outs

    def reset(self):
        return self._fanout(lambda g, env: env.reset())

    def step(self, actions=None):
        """actions: list of per-group tensors (or None).  Returns a list of (obs, rew, done) per group; the caller's
        current stream is ordered after every group's step."""
        return self._fanout(lambda g, env: env.step(None if actions is None else actions[g]))

    def capture_steps(self, n_steps, actions=None):
        """One CUDA graph of n_steps rounds over all groups (fork / join on the capturing stream)."""
        t = self.torch
        cap = t.cuda.Stream(device=self.device)
        cap.wait_stream(t.cuda.current_stream(self.device))
        with t.cuda.stream(cap):
            self.step(actions)                                   # warm-up outside the capture
        cap.synchronize()
        g = t.cuda.CUDAGraph()
        with t.cuda.graph(g, stream=cap):
            for g_i, (env, st) in enumerate(zip(self.envs, self.streams)):
                st.wait_stream(cap)
                with t.cuda.stream(st):
                    for _ in range(int(n_steps)):
                        env.step(None if actions is None else actions[g_i])
            for st in self.streams:
                cap.wait_stream(st)
        return g

    def stats(self):
        import numpy as np
        return np.sum([e.stats() for e in self.envs], axis=0)

    def close(self):
        for e in self.envs:
            e.close()

"""Where the e2e step (qs_step_host, 65,536 envs, pinned host buffers) spends its time:
  * zero-copy as shipped (the kernel reads the actions and posts the results across PCIe);
  * no action input at all (in-kernel Mellinger controller): what the posted result writes cost by themselves.
(tools/micro/pcie_write_bench.cu measures the link without any simulation work.  A hybrid - copy engine brings the actions in
1 / 2 / 4 chunks, the chunk launches post the results - was tried in round 2: 137 - 151 us against 124 - 126, dropped.)
python tools/e2e_probe.py [steps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402


def run(label, cfg, env_knobs, steps, with_actions=True):
    for k in ("QUADSIM_HOST_CTAS",):
        os.environ.pop(k, None)
    os.environ.update(env_knobs)
    n, dev = bench.N_BATCH, torch.device("cuda", 0)
    a_h = (torch.rand((n, 4)) * 2 - 1).pin_memory() if with_actions else None
    o_h, r_h = torch.empty((n, 18)).pin_memory(), torch.empty((n,)).pin_memory()
    d_h = torch.empty((n,), dtype=torch.uint8).pin_memory()
    e = bench.make_env(n, dev, seed=1, cfg=cfg)
    e.reset()
    for _ in range(20):
        e.step_host(a_h, o_h, r_h, d_h)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        e.step_host(a_h, o_h, r_h, d_h)
    torch.cuda.synchronize()
    us = (time.perf_counter() - t0) / steps * 1e6
    nbytes = n * (77 + (16 if with_actions else 0))
    print("%-46s %7.1f us/step  %.3e env-steps/s  %5.1f GB/s over the link  [%s]" % (label, us, n / us * 1e6, nbytes / us * 1e-3, e.host_path), flush=True)
    e.close()


if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1000
    for rep in range(2):
        run("zero-copy (shipped)", "cfg2", {}, steps)
        run("zero-copy, no action input (cfg5 mellinger)", "cfg5", {}, steps, with_actions=False)
        run("zero-copy, no action input, host_ctas 0", "cfg5", {"QUADSIM_HOST_CTAS": "0"}, steps, with_actions=False)

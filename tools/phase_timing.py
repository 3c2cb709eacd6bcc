"""Experiment: where does a one-wave 65,536-env step spend its time?  Needs a library built with -DQS_PHASE_TIMING
(tools/build_timing.sh -> gym_art_b200/_exp/libquadsim_timing.so, selected with QUADSIM_LIB).  Every warp records
%globaltimer at kernel entry (after the PDL wait), when its tile's data has arrived, when the step is computed and
when its stores are issued."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402
import torch  # noqa: E402

n = 65536
warps = n // 32
buf = torch.zeros((warps, 5), dtype=torch.int64, device="cuda")
os.environ["QUADSIM_TIMING_PTR"] = str(buf.data_ptr())
import bench  # noqa: E402

dev = torch.device("cuda", 0)
ring = 16
envs, acts = [], []
for b in range(ring):
    e = bench.make_env(n, dev, seed=1, env_id_offset=b * n)
    e.reset()
    envs.append(e)
    acts.append(torch.rand((n, 4), device=dev) * 2 - 1)
bench.RING_PERIOD[0] = ring
g = torch.cuda.CUDAGraph()
st = torch.cuda.Stream()
with torch.cuda.stream(st):
    for i in range(64):
        envs[i % ring].step(acts[i % ring])
st.synchronize()
with torch.cuda.graph(g, stream=st):
    for i in range(160):
        envs[i % ring].step(acts[i % ring])
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
t = buf.cpu().numpy().astype(np.float64)           # stamps of the LAST step of the graph
t0, t1, t2, t3, sm = t[:, 0], t[:, 1], t[:, 2], t[:, 3], t[:, 4].astype(int)
base = t0.min()
q = lambda x: "min %.0f  p50 %.0f  p90 %.0f  max %.0f" % (x.min(), np.median(x), np.quantile(x, 0.9), x.max())
print("ns since the first warp started:")
print("  kernel entry      ", q(t0 - base))
print("  data arrived      ", q(t1 - base))
print("  step computed     ", q(t2 - base))
print("  stores issued     ", q(t3 - base))
print("per-warp phases (ns): load %s | compute %s | flush %s" % (q(t1 - t0), q(t2 - t1), q(t3 - t2)))
cnt = np.bincount(sm, minlength=148) // 4
print("CTAs per SM: min %d max %d;  last-store time on 4-CTA SMs %.0f vs 3-CTA SMs %.0f" % (
    cnt.min(), cnt.max(), np.mean([t3[sm == k].max() - base for k in range(148) if cnt[k] == 4]),
    np.mean([t3[sm == k].max() - base for k in range(148) if cnt[k] == 3] or [0])))

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
buf = torch.zeros((2, warps, 6), dtype=torch.int64, device="cuda")
os.environ["QUADSIM_TIMING_PTR"] = str(buf.data_ptr())
os.environ["QUADSIM_TIMING_RING"] = sys.argv[1] if len(sys.argv) > 1 else "16"
import bench  # noqa: E402

dev = torch.device("cuda", 0)
ring = int(sys.argv[1]) if len(sys.argv) > 1 else 16
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
NSTEP = 160
with torch.cuda.graph(g, stream=st):
    for i in range(NSTEP):
        envs[i % ring].step(acts[i % ring])
for _ in range(3):
    g.replay()
torch.cuda.synchronize()
# the two record sets hold the last launch of the handles whose step parity is 0 / 1; the last two launches of the graph
# belong to batches ring-2 and ring-1, each stepped (64 + 4 * 160) / 16 = 44 times -> both have the same parity: take
# the one batch that was launched last per parity slot.  Simpler: look at both sets and order them by entry time.
t = buf.cpu().numpy().astype(np.float64)
sets = sorted([t[0], t[1]], key=lambda x: x[:, 0].min())
q = lambda x: "min %6.0f  p10 %6.0f  p50 %6.0f  p90 %6.0f  max %6.0f" % (x.min(), np.quantile(x, 0.1), np.median(x), np.quantile(x, 0.9), x.max())
for name, r in zip(("earlier launch", "later launch"), sets):
    if r[:, 0].min() == 0:
        continue
    base = r[:, 1].min()                                # first warp released from the wait
    print("%s - ns relative to the first warp passing griddepcontrol.wait:" % name)
    print("  CTA entry (prologue)", q(r[:, 0] - base))
    print("  wait passed         ", q(r[:, 1] - base))
    print("  data arrived        ", q(r[:, 2] - base))
    print("  step computed       ", q(r[:, 3] - base))
    print("  stores issued       ", q(r[:, 4] - base))
    print("  per-warp phases: prologue->wait %s | load %s | compute %s | flush %s" % (
        q(r[:, 1] - r[:, 0]), q(r[:, 2] - r[:, 1]), q(r[:, 3] - r[:, 2]), q(r[:, 4] - r[:, 3])))
    sm = r[:, 5].astype(int)
    cnt = np.bincount(sm, minlength=148)
    print("  warps per SM: min %d max %d" % (cnt.min(), cnt.max()))
if sets[0][:, 0].min() > 0:
    print("later launch's first wait-pass minus earlier launch's last store: %.0f ns" % (sets[1][:, 1].min() - sets[0][:, 4].max()))
    print("later launch's last store minus earlier launch's last store (step time): %.0f ns" % (sets[1][:, 4].max() - sets[0][:, 4].max()))

"""DRAM traffic of the bench headline in the regime that is timed: one CUDA graph of `steps` qs_step launches cycling
through the ring of 16 x 65,536-env cfg2 batches (366 MB > L2), replayed a few times.  Run it under

    ncu --graph-profiling graph --profile-from-start off --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum \
        --clock-control none --csv --log-file gpurun_out/ring_traffic.csv python tools/ring_traffic.py

Each profiled workload is then one whole graph launch; bytes / steps = steady-state DRAM traffic per step launch (the last
replays, with the ring warm).  Without ncu it prints the plain step time of the same graph."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == "__main__":
    steps = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    n, ring = bench.N_BATCH, 16
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev)
    gen.manual_seed(1)
    envs, acts = [], []
    for b in range(ring):
        e = bench.make_env(n, dev, seed=1, env_id_offset=b * n)
        e.reset()
        envs.append(e)
        acts.append(torch.rand((n, 4), device=dev, generator=gen) * 2 - 1)
    stream = torch.cuda.Stream(device=dev)
    with torch.cuda.stream(stream):
        for i in range(2 * ring):
            envs[i % ring].step(acts[i % ring])
    stream.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=stream):
        for i in range(steps):
            envs[i % ring].step(acts[i % ring])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for rep in range(6):
        if rep == 3:
            torch.cuda.profiler.start()                      # ncu --profile-from-start off: only the warm replays are profiled
        with torch.cuda.stream(stream):
            e0.record(stream)
            g.replay()
            e1.record(stream)
        stream.synchronize()
        print("replay %d: %d steps, %.3f us per step (idle-to-idle)" % (rep, steps, e0.elapsed_time(e1) * 1e3 / steps), flush=True)
    torch.cuda.profiler.stop()
    for e in envs:
        e.check_finite()

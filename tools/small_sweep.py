import sys
sys.path.insert(0, "/root/repo")
import torch, bench
dev = torch.device("cuda", 0)
for n in (1024, 4096, 8192, 12288, 16384, 24576, 32768, 40960, 49152, 65536, 75776):
    ring = 16
    envs, acts = [], []
    for b in range(ring):
        e = bench.make_env(n, dev, seed=1, env_id_offset=b * n)
        e.reset(); envs.append(e)
        acts.append(torch.rand((n, 4), device=dev) * 2 - 1)
    bench.RING_PERIOD[0] = ring
    K = 4000
    ms = bench.time_graph(torch, lambda i: envs[i % ring].step(acts[i % ring]), K, 64, dev, chunk=K)[0][0]
    print("n=%6d  %.3f us/step" % (n, ms * 1e3 / K), flush=True)
    for e in envs: e.close()

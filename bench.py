#!/usr/bin/env python
"""bench.py - env-steps/s of the fused quadrotor env-step path on B200, with roofline and CPU baseline.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N --steps K --warmup W

A "step" is one pass of the hot path (controller -> sim_steps x step1 -> reward/done -> auto-reset -> obs) over one
batch of envs.  Workload (BASELINE.json configs[1], "cfg2"): Hummingbird, RawControl zero-middle, no sensor noise,
OU thrust noise on (model default), obs xyz_vxyz_R_omega, 65,536 envs per batch, auto-reset, fp32, Philox noise.
A ring of independent 65,536-env batches whose combined footprint exceeds the 126 MB L2 is cycled so that every
step finds its inputs in HBM, not in L2.  The same run also times the HBM-resident size (4,194,304 envs per
step, "hbm_resident" in the JSON) on which BASELINE.md quotes the roofline fraction.

One JSON line is printed by rank 0 (contract in the task statement): value = whole-job env-steps/s on device-resident
inputs; e2e = the same metric through qs_step_host with pinned HOST buffers (H2D + D2H inside the timed region);
roofline = algorithmic bytes / measured launch duration vs MEASURED_PEAKS.json; cpu_baseline = the NumPy oracle port
of the reference timed on this box's host cores.
"""
import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_BATCH = 65536              # cfg2
N_LARGE = 4 * 1024 * 1024    # HBM-resident size the roofline fraction is quoted on
SIM_STEPS = 2
EP_TIME = 5                  # ep_len 500 -> 501-step episodes
BYTES_PER_ENV_STEP = 2 * 4 * 32 + 4 * 4 + 4 * 18 + 4 + 1   # 349 B: S=32 state words r+w, 4 action words, 18 obs, reward, done
METRIC = "env-steps/sec at 1/2/4/8 B200 (% HBM roofline) vs host-core numpy ref"
WORKLOAD = ("cfg2: Hummingbird (DefaultQuad) RawControl zero-middle, no sensor noise, OU thrust noise on, "
            "obs xyz_vxyz_R_omega, sim_steps=2, ep_len=500, auto-reset, %d envs per step" % N_BATCH)


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        try:
            return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


# ----------------------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ----------------------------------------------------------------------------------------------------------------
class ClockSampler(object):
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.samples = []
        self.proc = None
        self.index = index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                                          "--format=csv,noheader,nounits", "-lms", "100"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.samples.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], None, set()
        for s in self.samples:
            p = [x.strip() for x in s.split(",")]
            if len(p) < 7:
                continue
            try:
                sm.append(float(p[0]))
                smax = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        # median over the busiest half of the samples (the sampler also sees the idle gaps around the timed region)
        busy = sm[len(sm) // 2:] if sm else []
        med = busy[len(busy) // 2] if busy else None
        return {"sm_mhz": med, "sm_max_mhz": smax, "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------------
# CPU baseline: the NumPy oracle port of the reference, one env per process (the reference's own object model)
# ----------------------------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seconds, seed, n_envs = args[:3]
    cfg = args[3] if len(args) > 3 else "cfg2"
    import numpy as np
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    from oracle import quadsim_oracle as qo
    params = _hummingbird_params()
    if cfg == "cfg1":       # BASELINE configs[0]: ./quadrotor.py defaults - Mellinger closed loop, default sensor noise
        env = qo.OracleVecEnv(params, n_envs, controller=qo.CTRL_MELLINGER, ep_time=EP_TIME, sim_steps=SIM_STEPS, auto_reset=True,
                              sense_noise="default", obs_repr="xyz_vxyz_R_omega_act")
    else:
        env = qo.OracleVecEnv(params, n_envs, controller=qo.CTRL_RAW, ep_time=EP_TIME, sim_steps=SIM_STEPS, auto_reset=True)
    rng = np.random.RandomState(seed)
    sense = cfg == "cfg1"

    def one_step():
        env.step(rng.uniform(-1, 1, (n_envs, 4)), ou_z=rng.standard_normal((n_envs, SIM_STEPS, 4)),
                 obs_z=rng.standard_normal((n_envs, qo.OBS_Z)) if sense else None,
                 reset_z=rng.uniform(-1, 1, (n_envs, qo.RESET_Z)))

    env.reset(reset_z=rng.uniform(-1, 1, (n_envs, qo.RESET_Z)))
    for _ in range(20):
        one_step()
    steps = 0
    t0 = time.perf_counter()
    while True:
        for _ in range(50):
            one_step()
        steps += 50
        dt = time.perf_counter() - t0
        if dt >= seconds:
            return steps * n_envs / dt


def _hummingbird_params():
    # parameter *data* for the oracle (the oracle derives its own constants from it)
    from gym_art_b200.quadrotor import quad_models
    return quad_models.defaultquad_params()


# Single-env steps/s measured in the BUILD container (8 cores, same interpreter, one process; tools/port_vs_live.py): the
# live reference under the import stubs vs the oracle port this arm times.  /root/reference does not exist on the GPU box,
# so the port is what runs there; per single env it is somewhat SLOWER than the reference's own Python step, i.e. the
# CPU figures of this file understate the live reference by the factor below (it is 1e5 away from mattering).
PORT_VS_LIVE = {"cfg1": {"live_reference_steps_per_s": 1521.5, "oracle_port_steps_per_s": 1020.3},
                "cfg2": {"live_reference_steps_per_s": 3196.9, "oracle_port_steps_per_s": 2797.3},
                "live_over_port": {"cfg1": 1.49, "cfg2": 1.14}, "where": "build container (8 cores), tools/port_vs_live.py"}


def cpu_baseline(seconds=6.0, cores=None):
    cores = cores or os.cpu_count() or 1
    ctx = mp.get_context("spawn")
    single = _cpu_worker((min(seconds, 4.0), 0, 1))
    cfg1 = _cpu_worker((min(seconds, 4.0), 1, 1, "cfg1"))
    with ctx.Pool(cores) as pool:
        rates = pool.map(_cpu_worker, [(seconds, 100 + k, 1) for k in range(cores)])
        batched = pool.map(_cpu_worker, [(seconds / 2, 200 + k, 4096) for k in range(cores)])
    return {
        "value": float(sum(rates)), "unit": "env-steps/s", "cores": cores, "kind": "port",
        "sample": "NumPy oracle port (oracle/quadsim_oracle.py, bit-verified against the live reference), cfg2 semantics, "
                  "one env per process as in the reference, %d processes x %.0f s, sum of per-process rates" % (cores, seconds),
        "single_env_1core": float(single),
        "cfg1_single_env_1core": float(cfg1),
        "cfg1_note": "BASELINE configs[0] (Hummingbird, Mellinger closed loop, default SensorNoise, obs xyz_vxyz_R_omega_act, "
                     "ep_time 5, sim_steps 2), one env, one core, same port",
        "port_vs_live_reference": PORT_VS_LIVE,
        "batched_numpy_all_cores": float(sum(batched)),
        "batched_note": "same port vectorised over 4096 envs per process (not how the reference runs; upper bound for NumPy)",
    }


# ----------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------
def make_env(n, device, seed, env_id_offset=0, cfg="cfg2"):
    from gym_art_b200 import QuadVecEnv
    kw = dict(obs_repr="xyz_vxyz_R_omega", ep_time=EP_TIME, sim_steps=SIM_STEPS, device=device, dtype="float32", seed=seed,
              noise_mode="philox", auto_reset=True, env_id_offset=env_id_offset)
    if cfg == "cfg2":      # Hummingbird RawControl, no sensor noise
        return QuadVecEnv(n, dynamics_params="DefaultQuad", controller="raw", raw_control_zero_middle=True, sense_noise=None, **kw)
    if cfg == "cfg3":      # CrazyFlie OmegaThrustControl + SensorNoise + random init
        return QuadVecEnv(n, dynamics_params="Crazyflie", controller="omega_thrust", sense_noise="default",
                          init_random_state=True, **kw)
    if cfg == "cfg4":      # swarms of 8 Hummingbirds on a ring formation, pairwise contact
        from gym_art_b200.quadrotor_multi.quadrotor_multi import ring_goals
        return QuadVecEnv(n, dynamics_params="DefaultQuad", controller="raw", n_agents=8, formation=ring_goals(8), **kw)
    if cfg == "cfg5":      # Hummingbird Mellinger closed loop (no action input)
        return QuadVecEnv(n, dynamics_params="DefaultQuad", controller="mellinger", **kw)
    raise ValueError(cfg)


# (config, envs per GPU, algorithmic bytes per env step): the other BASELINE.json configs, timed briefly for context
OTHER = (("cfg3", 1 << 20, 349), ("cfg4", 262144 * 8, 349), ("cfg5", 8 << 20, 333))
# dram__bytes_read.sum + dram__bytes_write.sum per launch from committed ncu captures.  The headline size is taken in the regime
# that is timed - ncu --graph-profiling graph over warm replays of a 64-step graph cycling through the ring (tools/ring_traffic.py):
# 21.9 MB per launch, 0.96 of the 22.87 MB algorithmic; the large size from the --set full capture of one launch.
NCU_TRAFFIC = {N_BATCH: 21.89e6, N_LARGE: 1408.0e6}
NCU_TRAFFIC_SOURCE = ("profiles/r2_ring_traffic.txt (ncu --graph-profiling graph over warm ring replays: read 9.43 MB + write 12.47 MB per "
                      "65,536-env launch)")


def time_graph(torch, fn_steps, n_steps, warmup, device, dist=None, reducer=None, chunk=501, repeats=1, launches_of=None,
               start=0, lead_in=0):
    """Captures `fn_steps(i)` for i in a chunk into CUDA graphs and times exactly n_steps steps with CUDA events on
    the launching stream, `repeats` times (each repeat bracketed by barrier + synchronize).  `reducer` (an
    EpisodeStatsReducer) is told after every chunk how many steps have been enqueued since the reset; it fires the
    path's only collective on the episode cadence, on its own stream.  Returns (list of elapsed ms per repeat on this
    rank, kernel launches inside one timed repeat counted through qs_launch_count at capture time)."""
    stream = torch.cuda.Stream(device=device)
    graphs, graph_launches = {}, {}
    # lead_in > 0 (only when the K steps fit one graph): the graph holds `lead_in` untimed steps followed by the K timed
    # steps, and the two timing events are event-record NODES of that graph - e0 on a forked branch behind the last
    # lead-in step (so the chain of dependent launches is not interrupted), e1 behind the K-th timed step.  The interval
    # is then completion of the step before the timed window -> completion of its K-th step: exactly K steps with the
    # launch pipeline in its steady state, independent of K.  Without it a 20-step region also measures the idle-to-busy
    # ramp of one graph launch (about 5 us, i.e. 4 % of 20 steps).
    lead = int(lead_in) if n_steps <= chunk else 0
    side = torch.cuda.Stream(device=device) if lead else None
    ge0 = torch.cuda.Event(enable_timing=True, external=True) if lead else None
    ge1 = torch.cuda.Event(enable_timing=True, external=True) if lead else None

    def graph_for(first, count):
        key = (first % RING_PERIOD[0], count)
        if key not in graphs:
            before = launches_of() if launches_of else 0
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=stream):
                if lead:
                    for i in range(first - lead, first):
                        fn_steps(i)
                    side.wait_stream(stream)
                    ge0.record(side)
                for i in range(first, first + count):
                    fn_steps(i)
                if lead:
                    ge1.record(stream)
                    stream.wait_stream(side)
            graphs[key] = g
            graph_launches[key] = ((launches_of() - before) if launches_of else count + lead) - lead
        return key

    with torch.cuda.stream(stream):
        for i in range(start, start + warmup):
            fn_steps(i)
    stream.synchronize()
    enq = [start + warmup]                           # steps enqueued since the reset

    def run_plan(plan):
        for key in plan:
            graphs[key].replay()
            enq[0] += key[1] + lead
            if reducer is not None:
                reducer.after_steps(enq[0], stream)

    def make_plan(first):
        plan, done = [], 0
        while done < n_steps:
            c = min(chunk, n_steps - done)
            plan.append(graph_for(first + done, c))
            done += c
        return plan

    first0 = enq[0] + lead
    plan = make_plan(first0)
    with torch.cuda.stream(stream):                  # one untimed pass (graph instantiation / upload stay outside)
        run_plan(plan)
    stream.synchronize()
    out = []
    for _ in range(repeats):
        with torch.cuda.stream(stream):              # untimed eager steps up to the ring phase the graphs were captured at
            while (enq[0] + lead - first0) % RING_PERIOD[0]:
                fn_steps(enq[0])
                enq[0] += 1
        plan = make_plan(enq[0] + lead)
        stream.synchronize()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            e0.record(stream)
            run_plan(plan)
            e1.record(stream)
        stream.synchronize()
        torch.cuda.synchronize(device)
        out.append(ge0.elapsed_time(ge1) if lead else e0.elapsed_time(e1))
    return out, sum(graph_launches[k] for k in plan)


def reduce_max(torch, dist, device, values):
    """max over ranks of every entry (a rank's own CUDA-event time); identity without a process group"""
    t = torch.tensor(list(values), device=device, dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def median(xs):
    s = sorted(xs)
    return s[len(s) // 2] if len(s) % 2 else 0.5 * (s[len(s) // 2 - 1] + s[len(s) // 2])


RING_PERIOD = [1]


def _cpus_of(spec):
    cpus = set()
    for part in spec.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def bind_to_gpu_numa_node(torch, local):
    """Pin this rank's threads (and so its first-touch pinned host buffers) to the CPUs next to its GPU.  Sources, in
    order: sysfs numa_node of the GPU's PCI function; `nvidia-smi topo -m` CPU / NUMA affinity columns (what the
    driver's own topology probe prints).  Returns {"node": .., "cpus": n, "source": ..} - never silently nothing."""
    info = {"node": None, "cpus": None, "source": "none"}
    cpus = None
    try:
        pr = torch.cuda.get_device_properties(local)
        bdf = "%04x:%02x:%02x.0" % (pr.pci_domain_id, pr.pci_bus_id, pr.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bdf).read())
        if node >= 0:
            cpus = _cpus_of(open("/sys/devices/system/node/node%d/cpulist" % node).read())
            info.update(node=node, source="sysfs")
    except Exception:  # noqa: BLE001
        pass
    if cpus is None:
        try:
            import re
            out = subprocess.run(["nvidia-smi", "topo", "-m"], capture_output=True, text=True, timeout=20).stdout
            out = re.sub(r"\x1b\[[0-9;]*m", "", out)
            for ln in out.splitlines():
                cols = ln.split("\t")
                if cols and cols[0].strip() == "GPU%d" % local:
                    aff = [c.strip() for c in cols[1:] if re.fullmatch(r"\d+(-\d+)?(,\d+(-\d+)?)*", c.strip())]
                    if aff:
                        cpus = _cpus_of(aff[0])
                        info.update(node=int(aff[1]) if len(aff) > 1 and aff[1].isdigit() else None, source="nvidia-smi topo")
                    break
        except Exception:  # noqa: BLE001
            pass
    try:
        if cpus:
            cpus &= os.sched_getaffinity(0)
            if cpus:
                os.sched_setaffinity(0, cpus)
                info["cpus"] = len(cpus)
    except Exception:  # noqa: BLE001
        pass
    return info


def run_ours(args):
    import torch
    import torch.distributed as dist
    from gym_art_b200.stats import EpisodeStatsReducer
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    multi = world > 1
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    device = torch.device("cuda", local)
    torch.cuda.set_device(device)
    numa = bind_to_gpu_numa_node(torch, local)        # before any pinned allocation (first touch)
    if multi:
        dist.init_process_group("nccl", device_id=device)
    pg = dist if multi else None
    hbm_peak, peak_src = load_peaks()
    K, W = args.steps, max(args.warmup, 3)
    EP_STEPS = int(EP_TIME * 200 / SIM_STEPS) + 1     # 501 steps per episode (ep_len + 1, quadrotor.py:986-987)

    # ---- primary workload: ring of 65,536-env batches (> L2) ---------------------------------------------------
    ring = args.ring
    RING_PERIOD[0] = ring
    envs, acts = [], []
    gen = torch.Generator(device=device)
    gen.manual_seed(1 + rank)
    for b in range(ring):
        e = make_env(N_BATCH, device, seed=1, env_id_offset=(rank * ring + b) * N_BATCH)
        e.reset()
        envs.append(e)
        acts.append(torch.rand((N_BATCH, 4), device=device, generator=gen) * 2 - 1)
    ring_bytes = ring * (envs[0].state.numel() * 4 + N_BATCH * (16 + 72 + 4 + 1))

    def step_ring(i):
        b = i % ring
        envs[b].step(acts[b])

    def ring_launches():
        return sum(e.launch_count for e in envs)

    # the path's only collective: all-reduce of {sum ret, sum ret^2, episodes} (NCCL), fired whenever the step count
    # crosses a multiple of the episode length, on a side stream (gym_art_b200/stats.py).  With a ring of R batches a
    # batch ends an episode every R x 501 bench steps, so this cadence is R times denser than the episodes need.
    reducer = EpisodeStatsReducer([e.stats_tensor() for e in envs], EP_STEPS)

    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    # short timed regions (the driver's --steps 20 is 140 us) are repeated and the median reported; all repeats listed
    repeats = max(1, min(9, 4000 // max(K, 1)))
    ms_all, launches = time_graph(torch, step_ring, K, W, device, pg, reducer, chunk=EP_STEPS, repeats=repeats,
                                  launches_of=ring_launches, lead_in=ring)
    ms_all = reduce_max(torch, pg, device, ms_all)            # per repeat: max over ranks
    ms_max = median(ms_all)
    # for the record: the same K steps timed from an idle GPU (events around one graph replay, no lead-in steps)
    cold = None
    if K <= EP_STEPS:
        ms_cold, _ = time_graph(torch, step_ring, K, 0, device, pg, None, chunk=EP_STEPS, repeats=repeats, start=1 << 20)
        cold = median(reduce_max(torch, pg, device, ms_cold)) / K
    value = world * N_BATCH * K / (ms_max * 1e-3)
    step_us = ms_max * 1e3 / K
    achieved = N_BATCH * BYTES_PER_ENV_STEP / (step_us * 1e-6) / 1e9
    collectives_in_timed_runs = reducer.fired
    # one reduction at the end of the run (untimed) so that the job-wide statistics are reported whatever K was
    reducer.reduce_now(torch.cuda.current_stream(device))
    ep_stats = reducer.result()
    # cost of one such collective, measured on its own (CUDA events on the side stream, 50 back-to-back reductions)
    collective_us = None
    if multi:
        cs = reducer.stream
        buf = torch.zeros((ring, 3), dtype=torch.float64, device=device)
        with torch.cuda.stream(cs):
            for _ in range(5):
                dist.all_reduce(buf)
            cs.synchronize()
            dist.barrier()
            c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            c0.record(cs)
            for _ in range(50):
                dist.all_reduce(buf)
            c1.record(cs)
            cs.synchronize()
        collective_us = reduce_max(torch, pg, device, [c0.elapsed_time(c1) * 1e3 / 50])[0]

    # ---- context only: the same 65,536-env batch stepped back to back (what a training loop with one batch does):
    # its 23 MB working set then lives in the 126 MB L2 and the step never waits for HBM
    RING_PERIOD[0] = 1
    KR = max(200, min(K, 20040))
    ms_r, _ = time_graph(torch, lambda i: envs[0].step(acts[0]), KR, W, device, pg)
    ms_r = reduce_max(torch, pg, device, ms_r)[0]
    l2_resident = {"n_envs_per_gpu_step": N_BATCH, "steps": KR, "ms_per_step": ms_r / KR,
                   "value": world * N_BATCH * KR / (ms_r * 1e-3), "unit": "env-steps/s",
                   "note": "NOT the headline: one batch re-stepped, inputs L2-resident (no flush between iterations)"}

    # ---- HBM-resident size ----------------------------------------------------------------------------------------
    for e in envs[1:]:
        e.close()
    big = make_env(N_LARGE, device, seed=2, env_id_offset=rank * N_LARGE)
    big.reset()
    big_act = torch.rand((N_LARGE, 4), device=device, generator=gen) * 2 - 1
    KL = max(20, min(200, K // 10))
    ms_l, _ = time_graph(torch, lambda i: big.step(big_act), KL, 5, device, pg)
    ms_l = reduce_max(torch, pg, device, ms_l)[0]
    big_value = world * N_LARGE * KL / (ms_l * 1e-3)
    big_achieved = N_LARGE * BYTES_PER_ENV_STEP / (ms_l * 1e-3 / KL) / 1e9
    variant = big.kernel_variant
    big.check_finite()
    big.close()
    del big, big_act
    others = {}
    total64 = 64 << 20
    for cfg, n_o, bytes_o in OTHER + (("cfg5_64M_job", total64 // world, 333),):
        e = make_env(n_o, device, seed=3, env_id_offset=rank * n_o, cfg=cfg.split("_")[0])
        e.reset()
        a_o = None
        if not cfg.startswith("cfg5"):
            a_o = torch.rand((n_o, 4), device=device, generator=gen) * 2 - 1
            if cfg == "cfg3":
                a_o[:, 0] *= 0.2
        KO = 40 if n_o <= (8 << 20) else 12
        ms_o, _ = time_graph(torch, lambda i: e.step(a_o), KO, 5 if n_o <= (8 << 20) else 3, device, pg)
        ms_o = reduce_max(torch, pg, device, ms_o)[0]
        e.check_finite()
        others[cfg] = {"n_envs_per_gpu_step": n_o, "value": world * n_o * KO / (ms_o * 1e-3), "unit": "env-steps/s",
                       "ms_per_step": ms_o / KO, "kernel_variant": e.kernel_variant, "bytes_per_env_step": bytes_o,
                       "roofline_frac": n_o * bytes_o / (ms_o * 1e-3 / KO) / 1e9 / hbm_peak}
        if cfg != "cfg5_64M_job":
            # the same launch over one WHOLE episode (501 steps, the reset step included): trajectories evolve - under random
            # actions the drones of a swarm end up in contact on the floor - and the data-dependent branches cost more late in
            # an episode than in the first steps after a reset (tools/phase_of_episode.py)
            ms_ep, _ = time_graph(torch, lambda i: e.step(a_o), EP_STEPS, 0, device, pg, chunk=EP_STEPS)
            ms_ep = reduce_max(torch, pg, device, ms_ep)[0]
            e.check_finite()
            others[cfg]["whole_episode"] = {"steps": EP_STEPS, "ms_per_step": ms_ep / EP_STEPS,
                                            "value": world * n_o * EP_STEPS / (ms_ep * 1e-3),
                                            "roofline_frac": n_o * bytes_o / (ms_ep * 1e-3 / EP_STEPS) / 1e9 / hbm_peak}
        if cfg == "cfg5_64M_job":
            others[cfg].update(scaling="strong", n_envs_job=total64,
                               note="BASELINE configs[4]: ONE 64 M-env job sharded by env index over the GPUs of the run")
        e.close()
        del e, a_o
        torch.cuda.empty_cache()
    # f1: K env steps fused in one launch with the state held in registers - the closed-loop
    # lever once a single step is at the HBM roofline (Mellinger in-kernel, no action input)
    KF, n_f = 8, 4 << 20
    e = make_env(n_f, device, seed=4, env_id_offset=rank * n_f, cfg="cfg5")
    e.reset()
    ms_f, _ = time_graph(torch, lambda i: e.step_k(KF, None), 20, 3, device, pg)
    ms_f = reduce_max(torch, pg, device, ms_f)[0]
    e.check_finite()
    others["cfg5_fused_k8"] = {"n_envs_per_gpu_step": n_f, "k_steps_per_launch": KF, "value": world * n_f * KF * 20 / (ms_f * 1e-3),
                               "unit": "env-steps/s", "ms_per_launch": ms_f / 20, "kernel_variant": e.kernel_variant,
                               "note": "qs_step_k: HBM is touched once per K env steps, so the per-step HBM roofline does not apply"}
    e.close()
    del e
    clocks = sampler.stop() if rank == 0 else None

    # ---- e2e: host buffers through qs_step_host -------------------------------------------------------------------
    e0 = envs[0]
    a_h = torch.empty((N_BATCH, 4), dtype=torch.float32).pin_memory()
    a_h.copy_(acts[0].cpu())
    o_h = torch.empty((N_BATCH, 18), dtype=torch.float32).pin_memory()
    r_h = torch.empty((N_BATCH,), dtype=torch.float32).pin_memory()
    d_h = torch.empty((N_BATCH,), dtype=torch.uint8).pin_memory()
    KE = max(200, min(K, 2000))
    s_host = torch.cuda.Stream(device=device)         # the caller's stream the host-buffer steps are ordered on
    with torch.cuda.stream(s_host):
        for _ in range(5):
            e0.step_host(a_h, o_h, r_h, d_h)
        e2e_launch0 = e0.launch_count
        if multi:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        for _ in range(KE):
            e0.step_host(a_h, o_h, r_h, d_h)
        torch.cuda.synchronize(device)
    te = reduce_max(torch, pg, device, [time.perf_counter() - t0])[0]
    e2e_value = world * N_BATCH * KE / te
    e2e_launches = e0.launch_count - e2e_launch0
    assert bool(torch.isfinite(o_h).all())
    # denominator for the e2e number: the same bytes per step moved by the copy engines (cudaMemcpyAsync on pinned
    # buffers, H2D + D2H on two streams), all ranks at once, no kernel - what the host link of this box sustains
    s_in, s_out = torch.cuda.Stream(device=device), torch.cuda.Stream(device=device)
    a_d = torch.empty((N_BATCH, 4), dtype=torch.float32, device=device)

    def copy_round():
        with torch.cuda.stream(s_in):
            a_d.copy_(a_h, non_blocking=True)
        with torch.cuda.stream(s_out):
            o_h.copy_(e0.obs, non_blocking=True)
            r_h.copy_(e0.rew, non_blocking=True)
            d_h.copy_(e0._done_u8, non_blocking=True)

    for _ in range(5):
        copy_round()
    torch.cuda.synchronize(device)
    if multi:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(KE):
        copy_round()
        s_in.synchronize()
        s_out.synchronize()                               # one round trip per step, like the synchronous API
    tc = reduce_max(torch, pg, device, [time.perf_counter() - t0])[0]
    bytes_step = N_BATCH * (16 + 72 + 4 + 1)
    host_link = {"what": "copy-engine round trip of the same %d bytes per step (H2D actions + D2H obs/reward/done), pinned buffers, "
                         "all %d ranks concurrently, synchronised every step; no kernel" % (bytes_step, world),
                 "us_per_step": tc / KE * 1e6, "gbps_per_rank": bytes_step / (tc / KE) / 1e9,
                 "gbps_all_ranks": world * bytes_step / (tc / KE) / 1e9,
                 "e2e_us_per_step": te / KE * 1e6, "e2e_gbps_per_rank": bytes_step / (te / KE) / 1e9,
                 "e2e_over_copy_engine_ceiling": (tc / KE) / (te / KE)}

    if rank == 0:
        cpu = cpu_baseline(seconds=args.cpu_seconds) if (world == 1 and not args.no_cpu) else None
        line = {
            "metric": METRIC, "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms_max / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_envs_per_gpu_step": N_BATCH, "sim_steps": SIM_STEPS, "substeps_per_s": value * SIM_STEPS,
                       "l2": "ring of %d independent batches, %.0f MB footprint > 126 MB L2, cycled" % (ring, ring_bytes / 1e6),
                       "launch": "CUDA graph of qs_step launches (device-resident clock, programmatic dependent launch), "
                                 "one kernel per step",
                       "timing": ("CUDA events around exactly K steps, max over ranks; repeated %d times (each repeat bracketed by "
                                  "barrier + synchronize), median reported, every repeat listed in ms_repeats. " % repeats) +
                                 ("K fits one CUDA graph: the graph holds %d untimed lead-in steps and then the K timed steps, and the "
                                  "two events are event-record nodes of that graph (e0 behind the last lead-in step, e1 behind the "
                                  "K-th step), i.e. K steps completion-to-completion with the launch pipeline in steady state; the "
                                  "same K steps timed from an idle GPU are in ms_per_step_from_idle" % ring if K <= EP_STEPS else
                                  "K spans %d graph replays of <= %d steps, events recorded on the stream around them" % (-(-K // EP_STEPS), EP_STEPS)),
                       "kernel_variant": variant, "parallelism": "env-index sharding x%d, no data-path collective" % world},
            "ms_repeats": ms_all, "ms_per_step_from_idle": cold,
            "e2e": {"value": e2e_value, "unit": "env-steps/s", "h2d_bytes_per_step": N_BATCH * 16,
                    "d2h_bytes_per_step": N_BATCH * (72 + 4 + 1), "steps": KE, "gpu_launches": e2e_launches,
                    "api": "qs_step_host on pinned host tensors: the step kernel reads the actions and writes obs/reward/done "
                           "across PCIe itself (%s)" % e0.host_path, "cpu_binding": numa, "host_link": host_link},
            "gpu_launches": launches,
            "collective": {"what": "all-reduce(SUM) of {sum ret, sum ret^2, episodes} per batch, NCCL, side stream, fired when the step "
                                   "count crosses a multiple of %d (EpisodeStatsReducer)" % EP_STEPS,
                           "fired_in_timed_runs": collectives_in_timed_runs, "collective_us": collective_us,
                           "job_episode_stats": ep_stats},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s", "frac": achieved / hbm_peak,
                         "traffic": NCU_TRAFFIC[N_BATCH], "traffic_source": NCU_TRAFFIC_SOURCE,
                         "peak_source": peak_src, "bytes_per_env_step": BYTES_PER_ENV_STEP,
                         "note": "65,536-env launches are launch/latency-bound (23 MB per launch); see hbm_resident"},
            "hbm_resident": {"n_envs_per_gpu_step": N_LARGE, "steps": KL, "value": big_value, "unit": "env-steps/s",
                             "ms_per_step": ms_l / KL,
                             "roofline": {"bound": "hbm", "achieved": big_achieved, "peak": hbm_peak, "unit": "GB/s",
                                          "frac": big_achieved / hbm_peak, "traffic": NCU_TRAFFIC[N_LARGE],
                                          "traffic_source": "profiles/r2_ncu_cfg2_4194304.txt"}},
            "other_workloads": others,
            "l2_resident_same_batch": l2_resident,
            "clocks": clocks,
            "cpu_baseline": cpu,
        }
        emit(line)
    if multi:
        dist.barrier()
        dist.destroy_process_group()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    cores = os.cpu_count() or 1
    # each "step" of the reference arm is a bounded sample: every host core steps its own single env (the
    # reference's object model) for a slice of wall-clock; K and W only scale the sample, capped to ~30 s
    seconds = max(4.0, min(30.0, 0.01 * (K + W)))
    cpu = cpu_baseline(seconds=seconds, cores=cores)
    line = {
        "impl": "reference", "metric": METRIC, "value": cpu["value"], "unit": "env-steps/s", "n_gpus": args.gpus, "steps": K,
        "warmup": W, "ms_per_step": 1e3 / cpu["value"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "note": "reference CPU path: oracle port, one env per process, all host cores"},
        "cpu_baseline": cpu,
        "e2e": {"value": cpu["value"], "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


_REAL_STDOUT = [None]


def quiet_stdout():
    """stdout carries exactly ONE JSON line: whatever libraries print there (NCCL's version banner, for one) is sent to
    stderr by pointing fd 1 at fd 2 for the duration of the run; emit() writes the line to the real stdout."""
    sys.stdout.flush()
    _REAL_STDOUT[0] = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT[0] is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_REAL_STDOUT[0], data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100200)
    ap.add_argument("--warmup", type=int, default=100)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--ring", type=int, default=16)
    ap.add_argument("--cpu-seconds", type=float, default=6.0)
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    quiet_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

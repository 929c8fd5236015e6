#!/usr/bin/env python
"""Headline benchmark: BEV pooled samples/sec (6 cams, fwd+bwd), BASELINE.json config 2.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config config2]

One "step" = one batch (B=8 samples x 6 cameras, 128x352, D=41, C=64, 200x200x1 BEV) through the whole hot
path, forward + backward:
    camera prep + frustum geometry -> keys -> counting sort -> intervals   (lss_build_plan, 4 kernels)
    feature staging -> fused lift+splat forward                           (lss_feat_stage, lss_liftsplat_fwd)
    fused backward                                                         (lss_liftsplat_bwd)

What the JSON line holds (all timed on the device with CUDA events unless stated):
`value`       device-resident inputs, every step replayed as ONE CUDA graph of the C-ABI calls, rotating batch sets
              larger than L2, `--in-flight` (default 4) independent batches in flight, one stream each.  The block
              of `--steps` steps is repeated `repeats` times (barrier + synchronize around every block) and the
              MEDIAN block counts, so the driver's short runs (--steps 20) give the same number as long ones.
`serial`      the same steps strictly one after the other (one batch in flight).
`cached_plan` serial steps with the plan reused (evaluation with a fixed rig): staging + fwd + bwd.
`module_api`  what train.py calls: eager `model.get_voxels(x, rots, ...)` + backward on the reference's own LSS
              class with lss2_multimodal_nu_b200.patch installed, one stream, per-call allocations and all.
`reference_gpu` the UNMODIFIED reference's get_geometry + get_cam_feats + voxel_pooling (its PyTorch code) forward +
              backward on the same B200, same shapes: the "PyTorch-on-GPU" denominator of the north star.
`e2e`         pipeline.HostPipeline (the public API for fixed shapes) fed from pinned HOST buffers: H2D of
              depth/feat/calibration and D2H of the gradients every step, six steps in flight, host wall clock.
`roofline`    the dominant kernel (fused forward: it writes the whole BEV map); `roofline_bwd` the backward.
              Kernel time = the phase's launches for all batch sets captured back to back in one CUDA graph,
              replayed between two CUDA events on the launching stream; algorithmic bytes from SURVEY.md 8d.
`cpu_baseline` / `--impl reference`: the C restatement of the reference's algorithm (oracle/lss_oracle.c, "port")
              on the host cores.
Multi-GPU: the batch shards by sample, one process per GPU, no data-path collective (weak scaling); timing = max
over ranks.  Only the cpu_baseline / reference legs touch oracle/; the product path never does.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="config2")
    ap.add_argument("--sets", type=int, default=4, help="rotating buffer sets (working set > L2)")
    ap.add_argument("--in-flight", type=int, default=0,
                    help="independent batches in flight (one CUDA stream each); 0: LSS_BENCH_IN_FLIGHT or 4")
    ap.add_argument("--repeats", type=int, default=0, help="timed blocks of --steps steps (median reported); 0: auto")
    ap.add_argument("--no-graph", action="store_true")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-gpu", action="store_true")
    ap.add_argument("--quick", action="store_true", help="kernel numbers only (tools/run_variants.sh)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="0: max(steps, 200)")
    return ap.parse_args()


METRIC = "BEV pooled samples/sec (6 cams, fwd+bwd)"
UNIT = "samples/s"


def workload_name(cfg):
    return ("LSS voxel_pooling fwd+bwd, batch %d/GPU, %d cams %dx%d, D=%d, C=%d, %dx%dx%d BEV"
            % (cfg.B, cfg.N, cfg.final_dim[0], cfg.final_dim[1], cfg.D, cfg.C, *cfg.nx))


def base_config(cfg, world, sets=4):
    """The `config` both arms print, key for key (the driver compares them); what is specific to our arm's
    launch (streams, graphs, layout) goes under `run`."""
    alg = cfg.algorithmic_bytes()
    return {"workload": workload_name(cfg), "global_batch": cfg.B * world,
            "parallelism": "sample-sharded x%d, no collective" % world,
            "l2": "inputs larger than L2: %d rotating batch sets (%.0f MB BEV+dBEV each)" % (sets, 2 * alg["bev"] / 1e6)}


# ---------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on host cores
# ---------------------------------------------------------------------------
def cpu_port_run(cfg, steps, warmup, budget_s=None):
    """Time the C restatement of the reference's algorithm (geometry + lift + voxel_pooling
    fwd + bwd) on one batch of the workload with all host threads."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import numpy as np
    import c_oracle as CO
    import lss_oracle as O
    from lss2_multimodal_nu_b200 import synthetic as S
    cal = S.make_calibration(cfg); ft = S.make_features(cfg); dbev = S.make_dbev(cfg)
    us, vs, ds = O.frustum_axes(cfg.final_dim, cfg.downsample, cfg.dbound)
    dx, bx, nx = O.gen_dx_bx(cfg.xbound, cfg.ybound, cfg.zbound)
    # all host threads the process may use (torchrun exports OMP_NUM_THREADS=1: override it)
    try:
        avail = len(os.sched_getaffinity(0))
    except AttributeError:
        avail = os.cpu_count() or 1
    CO.set_threads(avail)
    cores = CO.threads()

    def one():
        geom = CO.geometry(us, vs, ds, **cal)
        return CO.step(ft["depth"], ft["feat"], geom, dbev, dx, bx, nx, cfg.B, cfg.N, mode=0)

    for _ in range(max(1, warmup)):
        one()
    times = []
    t_begin = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        one()
        times.append(time.perf_counter() - t0)
        if budget_s is not None and time.perf_counter() - t_begin > budget_s:
            break
    mean = float(np.mean(times))
    return {"samples_per_s": cfg.B / mean, "ms_per_step": mean * 1e3, "cores": cores,
            "steps": len(times)}


def run_reference_arm(args, cfg):
    rank = int(os.environ.get("RANK", "0"))
    world = max(int(os.environ.get("WORLD_SIZE", "1")), args.gpus)
    if rank != 0:
        return  # rank 0 alone runs the CPU arm
    # every step is one batch of the workload on all host cores (~50 ms): the requested steps / warm-up are
    # honoured up to a time budget that keeps the whole run within a few minutes
    r = cpu_port_run(cfg, max(1, args.steps), max(1, args.warmup), budget_s=150.0)
    stock = reference_torch_cpu_leg(cfg)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["samples_per_s"], "unit": UNIT,
        "n_gpus": args.gpus, "steps": args.steps, "steps_timed": r["steps"], "warmup": max(3, args.warmup),
        "ms_per_step": r["ms_per_step"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(cfg, world, args.sets),
        "cpu_baseline": {"value": r["samples_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                         "sample": "one batch of %d samples per step, %d steps: C restatement of the "
                                   "reference algorithm (oracle/lss_oracle.c), OpenMP" % (cfg.B, r["steps"])},
        "e2e": {"value": r["samples_per_s"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        # the slower of the two CPU baselines is reported beside the line's value, not as it
        "reference_torch_cpu": stock,
    }
    print(json.dumps(line))


def reference_torch_cpu_leg(cfg, steps=3):
    """The UNMODIFIED reference class (oracle/_ref through oracle/ref_import.py) on the host cores: LSS.get_voxels
    + backward with stock PyTorch CPU kernels, all threads.  Reported beside the C restatement, which is the
    faster (hence the conservative) CPU baseline and stays the reference arm's `value`."""
    try:
        import torch
        sys.path.insert(0, os.path.join(ROOT, "oracle"))
        import ref_import
        if not ref_import.available():
            return {"unavailable": "reference tree not staged (oracle/stage_ref.py)"}
        if cfg.C != 64:
            return {"unavailable": "the reference's LSS class hard-codes camC = 64 (src/model_baseline.py:25)"}
        from lss2_multimodal_nu_b200 import synthetic as S
        torch.manual_seed(0)
        m = ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).train()
        cal = [torch.from_numpy(v) for v in S.make_calibration(cfg, 1234).values()]
        gen = torch.Generator(); gen.manual_seed(99)
        x = torch.randn(cfg.B * cfg.N, 512, cfg.fH, cfg.fW, generator=gen).requires_grad_(True)
        nx = [int(v) for v in m.nx]
        dbev = torch.randn(cfg.B, cfg.C * nx[2], nx[0], nx[1], generator=gen)

        def one():
            x.grad = None
            m.get_voxels(x, *cal).backward(dbev)
        one()
        t0 = time.perf_counter()
        for _ in range(steps):
            one()
        dt = (time.perf_counter() - t0) / steps
        return {"value": cfg.B / dt, "unit": UNIT, "ms_per_step": dt * 1e3, "steps": steps,
                "threads": torch.get_num_threads(), "kind": "reference",
                "note": "UNMODIFIED reference LSS.get_voxels (src/model_baseline.py:128-133) forward + backward, "
                        "stock PyTorch on the host cores"}
    except Exception as e:  # noqa: BLE001
        return {"unavailable": str(e)[:160]}


# ---------------------------------------------------------------------------
# clocks / host topology
# ---------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock and throttle reasons through NVML while the GPU works."""

    def __init__(self, index):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml as N
            N.nvmlInit()
            self.N = N
            self.h = N.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = N.nvmlDeviceGetMaxClockInfo(self.h, N.NVML_CLOCK_SM)
        except Exception:
            self.N = None

    def sample(self):
        if self.N is None:
            return
        N = self.N
        try:
            self.samples.append(N.nvmlDeviceGetClockInfo(self.h, N.NVML_CLOCK_SM))
            try:
                r = N.nvmlDeviceGetCurrentClocksEventReasons(self.h)
            except Exception:
                r = N.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
            names = {"hw_slowdown": getattr(N, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                     "hw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                     "sw_thermal_slowdown": getattr(N, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                     "sw_power_cap": getattr(N, "nvmlClocksThrottleReasonSwPowerCap", 0x4)}
            for k, bit in names.items():
                if r & bit:
                    self.reasons.add(k)
        except Exception:
            pass

    def start(self, period=0.002):
        def loop():
            while not self._stop.is_set():
                self.sample()
                time.sleep(period)
        self._thr = threading.Thread(target=loop, daemon=True)
        self._thr.start()

    def stop(self):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        s = sorted(self.samples)
        med = s[len(s) // 2] if s else None
        return {"sm_mhz": med, "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons),
                "samples": len(s)}


def _parse_cpulist(text):
    cpus = set()
    for part in text.strip().split(","):
        if not part:
            continue
        lo, _, hi = part.partition("-")
        cpus.update(range(int(lo), int(hi or lo) + 1))
    return cpus


def gpu_numa_cpus(index):
    """CPUs of the NUMA node the GPU's PCIe root hangs on (sysfs), else the set NVML reports, else None."""
    info = {}
    try:
        import pynvml as N
        N.nvmlInit()
        h = N.nvmlDeviceGetHandleByIndex(index)
        bus = N.nvmlDeviceGetPciInfo(h).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        dev_path = "/sys/bus/pci/devices/%s:%s" % (dom[-4:].lower(), rest.lower())
        info["pci"] = bus
        node = int(open(dev_path + "/numa_node").read().strip())
        info["numa_node"] = node
        if node >= 0:
            cpus = _parse_cpulist(open("/sys/devices/system/node/node%d/cpulist" % node).read())
            return cpus, info
        words = N.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() or 64) // 64 + 1)
        cpus = {64 * i + b for i, w in enumerate(words) for b in range(64) if (int(w) >> b) & 1}
        info["nvml_affinity"] = len(cpus)
        return cpus, info
    except Exception as e:  # noqa: BLE001
        info["error"] = str(e)[:80]
        return None, info


def bind_near_gpu(local, world_local):
    """Pin this process (and therefore the pages of its pinned buffers, first touch) to CPUs next to its GPU.
    When the platform reports one CPU set for every GPU, the ranks of a node share it evenly instead of
    piling onto the same cores."""
    try:
        before = os.sched_getaffinity(0)
    except AttributeError:
        return None, {}
    cpus, info = gpu_numa_cpus(local)
    cpus = (cpus & before) if cpus else set(before)
    if not cpus:
        cpus = set(before)
    if world_local > 1:
        ordered = sorted(cpus)
        share = max(2, len(ordered) // world_local)
        mine = ordered[(local * share) % len(ordered):][:share] or ordered
        cpus = set(mine)
    try:
        os.sched_setaffinity(0, cpus)
        info["bound_cpus"] = len(cpus)
    except OSError:
        return None, info
    return before, info


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
def run_ours(args, cfg):
    import numpy as np
    import torch
    import torch.distributed as dist
    from lss2_multimodal_nu_b200 import _abi, functional as F, shard, synthetic as S

    rank, local, world = shard.env_world()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; there is no CPU fallback for the product path")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    world_local = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
    all_cpus, host_info = bind_near_gpu(local, world_local)   # pinned buffers + the submitting thread next to the GPU
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _abi.load()

    # ---- per-rank inputs: `sets` independent batches so the working set exceeds L2 -------
    from lss2_multimodal_nu_b200.pipeline import LiftSplatStep, HostPipeline
    grid = F.GridSpec.from_bounds(cfg.xbound, cfg.ybound, cfg.zbound)
    us, vs, ds = F.frustum_axes(F.make_frustum(cfg.final_dim, cfg.downsample, cfg.dbound).to(dev))
    C = cfg.C
    stream = torch.cuda.Stream(dev)                       # timing / per-kernel stream
    in_flight = args.in_flight or int(os.environ.get("LSS_BENCH_IN_FLIGHT", "4"))
    in_flight = max(1, min(in_flight, args.sets))
    # batch set s runs on stream s % in_flight: consecutive batches overlap on the GPU, as independent
    # micro-batches do in a trainer (the path has no cross-batch dependency)
    lanes = [torch.cuda.Stream(dev) for _ in range(in_flight)]
    host, steps = [], []
    for s in range(args.sets):
        seed = shard.rank_seed(1234, rank, s)
        cal = S.make_calibration(cfg, seed); ft = S.make_features(cfg, seed)
        h = {k: torch.from_numpy(v).pin_memory() for k, v in {**cal, **ft}.items()}
        host.append(h)
        st_ = LiftSplatStep(cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, C, grid, us, vs, ds, device=dev,
                            capture=not args.no_graph, stream=lanes[s % in_flight])
        st_.load(h)
        with torch.cuda.stream(st_.stream):
            gen = torch.Generator(device=dev); gen.manual_seed(seed)
            st_._dbev.copy_(torch.randn(st_._dbev.shape, device=dev, generator=gen))
        steps.append(st_)
    torch.cuda.synchronize()
    KERNELS_PER_STEP = steps[0].kernels_per_step

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_block(n, use_lanes, runner):
        """n steps round-robin over the batch sets; CUDA-event time from before the first to after the last."""
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ln in use_lanes:
            ln.wait_stream(stream)
        for i in range(n):
            runner(steps[i % len(steps)])
        for ln in use_lanes:
            stream.wait_stream(ln)
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1)

    def median_of_blocks(n, repeats, use_lanes, runner):
        ms = []
        for _ in range(repeats):
            barrier()
            ms.append(timed_block(n, use_lanes, runner))
        barrier()
        ms.sort()
        return ms[len(ms) // 2], ms

    repeats = args.repeats or max(25, min(200, 4000 // max(1, args.steps)))
    run_default = lambda d: d.run()
    for i in range(max(3, args.warmup)):
        run_default(steps[i % len(steps)])
    barrier()
    # The timed block = exactly --steps steps, round-robin over the batch sets.  Each set's share of the block
    # is captured as ONE CUDA graph (its step, share times, back to back on its stream), so a block costs the
    # host len(sets) graph launches instead of --steps: the measurement is paced by the GPU, not by how fast
    # one Python thread can launch 50 us graphs on four streams (which is what cost scaling efficiency when 8
    # ranks shared a host).
    share = [len(range(s_, args.steps, len(steps))) for s_ in range(len(steps))]
    block_graphs = [d.capture_repeated(n) if (n > 0 and not args.no_graph) else None for d, n in zip(steps, share)]

    def run_block():
        if args.no_graph:
            for i in range(args.steps):
                steps[i % len(steps)].run()
            return
        for d, g in zip(steps, block_graphs):
            if g is not None:
                d.replay(g)

    def timed_steps_block():
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for ln in lanes:
            ln.wait_stream(stream)
        run_block()
        for ln in lanes:
            stream.wait_stream(ln)
        e1.record(stream)
        stream.synchronize()
        return e0.elapsed_time(e1)

    run_block(); barrier()
    clocks = ClockSampler(local)
    clocks.start()
    all_ms = []
    for _ in range(repeats):
        barrier()
        all_ms.append(timed_steps_block())
    barrier()
    clocks.stop()
    all_ms.sort()
    med_ms = all_ms[len(all_ms) // 2]
    value = shard.aggregate_throughput(cfg.B * args.steps, med_ms, dev)   # all samples / slowest rank
    med_ms = shard.max_over_ranks(med_ms, dev)
    ms_per_step = med_ms / args.steps

    # ---- the same steps strictly one after the other (one batch in flight) ----
    def on_lane0(fn):
        def run(d):
            with torch.cuda.stream(lanes[0]):
                fn(d)
        return run

    def serial_runner(d):
        if d._graph is not None:
            d._graph.replay()
        else:
            d._enqueue(lanes[0], d._side)

    n_serial = min(args.steps, 100)
    serial_ms, _ = median_of_blocks(n_serial, min(repeats, 15), [lanes[0]], on_lane0(serial_runner))
    serial_ms /= n_serial
    # steps of lane 0 only (their graphs were captured on that stream): sets 0, in_flight, 2*in_flight, ...
    own = [d for k, d in enumerate(steps) if k % in_flight == 0]
    saved = steps
    steps = own
    for d in own:
        d.run_cached_plan()
    torch.cuda.synchronize()
    cached_ms, _ = median_of_blocks(n_serial, min(repeats, 15), [lanes[0]], lambda d: d.run_cached_plan())
    cached_ms /= n_serial
    steps = saved

    # ---- per-phase timing: ONE CUDA graph holding the phase's launches for every batch set, back to back on
    # one stream (rotating sets: cold inputs), replayed between two CUDA events on that stream ----
    def time_kernel(name, replays):
        reps = 2
        with torch.cuda.stream(stream):
            for d in steps:                                    # warm (module load) outside the capture
                getattr(d, "enqueue_" + name)(stream.cuda_stream)
        stream.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=stream):
            cur = torch.cuda.current_stream(dev).cuda_stream
            for _ in range(reps):
                for d in steps:
                    getattr(d, "enqueue_" + name)(cur)
        ts = []
        with torch.cuda.stream(stream):
            g.replay()
            for _ in range(replays):
                a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
                a.record(stream); g.replay(); b.record(stream)
                ts.append((a, b))
        stream.synchronize()
        per = sorted(a.elapsed_time(b) * 1e3 / (reps * len(steps)) for a, b in ts)
        return {"mean_us": sum(per) / len(per), "median_us": per[len(per) // 2], "min_us": per[0]}

    kt = {k: time_kernel(k, 20) for k in ("plan", "stage", "fwd", "bwd")}

    # ---- roofline of the dominant kernel (fused forward) and of the backward --------------------
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        with open(peaks_path) as f:
            peak = float(json.load(f)["hbm_gbs"]); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    else:
        peak = 6650.0; peak_src = "fallback (B200_PROFILING.md)"
    alg = cfg.algorithmic_bytes()
    traffic = {}
    tr_path = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tr_path):
        try:
            with open(tr_path) as f:
                traffic = json.load(f).get(cfg.name, {})
        except Exception:
            traffic = {}

    def roof(kernel, key, name):
        us_ = kt[kernel]["median_us"]
        ach = alg[key] / (us_ * 1e-6) / 1e9
        return {"bound": "hbm", "kernel": name, "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                "traffic": traffic.get(kernel + "_dram_bytes"), "algorithmic_bytes_per_launch": alg[key],
                "kernel_us": us_, "peak_source": peak_src,
                "timing": "phase launches of all batch sets back to back in one CUDA graph, CUDA events on the "
                          "launching stream, median of 20 replays"}

    roofline = roof("fwd", "fwd", "pool_fwd_kernel<fused> (lss_liftsplat_fwd)")
    roofline["step_frac_of_hbm_roofline"] = (alg["total"] / (ms_per_step * 1e-3) / 1e9) / peak
    roofline["serial_step_frac_of_hbm_roofline"] = (alg["total"] / (serial_ms * 1e-3) / 1e9) / peak
    roofline["kernels_us"] = {k: round(v["median_us"], 2) for k, v in kt.items()}
    roofline_bwd = roof("bwd", "bwd", "liftsplat_bwd_kernel (lss_liftsplat_bwd)")

    if args.quick:
        if rank == 0:
            print(json.dumps({"value": value, "ms_per_step": ms_per_step, "roofline": roofline,
                              "roofline_bwd": roofline_bwd, "serial": {"ms_per_step": serial_ms},
                              "cached_plan": {"ms_per_step": cached_ms}}))
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- the drop-in as train.py calls it, and the stock reference on the same GPU -----------------
    module_api, reference_gpu = None, None
    if rank == 0 and not args.no_reference_gpu:
        module_api, reference_gpu = reference_legs(cfg, dev, steps[0])

    # ---- e2e: HostPipeline (public API), pinned host buffers, six steps in flight ----------
    e2e_steps = args.e2e_steps or max(args.steps, 200)
    pipe = HostPipeline(lambda: LiftSplatStep(cfg.B, cfg.N, cfg.D, cfg.fH, cfg.fW, C, grid, us, vs, ds,
                                              device=dev, capture=not args.no_graph), depth=6)
    for sl in pipe.slots:   # the upstream gradient is produced on the device by the downstream network
        with torch.cuda.stream(sl["step"].stream):
            sl["step"]._dbev.copy_(steps[0]._dbev)
    torch.cuda.synchronize()
    checksum = 0.0
    # the host-side "dataset": every slot's pinned input block holds one batch, laid out as the step
    # consumes it (a data loader writes there directly); each step copies it to the device
    host_blocks = [pipe.pack(host[k % len(host)], pipe.input_block(k)) for k in range(len(pipe.slots))]

    def e2e_run(n):
        nonlocal checksum
        for i in range(n):
            if pipe.in_flight() == len(pipe.slots):
                out = pipe.collect()
                checksum += float(out["d_depth"][0, 0, 0, 0])     # the host really reads the result
            pipe.submit()
        while pipe.in_flight():
            out = pipe.collect()
            checksum += float(out["d_depth"][0, 0, 0, 0])

    e2e_run(max(1000, 2 * e2e_steps))      # warm the host side too: the first few hundred steps of a fresh process run at half speed
    # the host link this number is bound by: pinned copies of a step's blocks, ALL ranks at once -- H2D alone,
    # then H2D and D2H together on two streams (what the pipeline does)
    blk = host_blocks[0]; st0 = pipe.slots[0]["step"]; dst = st0.in_block
    hout = pipe.slots[0]["h_out"]
    s_up, s_dn = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        dst.copy_(blk, non_blocking=True)
    torch.cuda.synchronize()
    link_gbs = 50 * blk.numel() * 4 / (time.perf_counter() - t0) / 1e9
    barrier()
    t0 = time.perf_counter()
    for _ in range(50):
        with torch.cuda.stream(s_up):
            dst.copy_(blk, non_blocking=True)
        with torch.cuda.stream(s_dn):
            hout.copy_(st0.out_block, non_blocking=True)
    torch.cuda.synchronize()
    bidir_gbs = 50 * (blk.numel() + hout.numel()) * 4 / (time.perf_counter() - t0) / 1e9
    link_all = shard.gather_floats(link_gbs, dev) if world > 1 else [link_gbs]
    bidir_all = shard.gather_floats(bidir_gbs, dev) if world > 1 else [bidir_gbs]
    reps = []
    for _ in range(9):                        # host wall clock is noisy: median of nine runs of e2e_steps
        barrier()
        t0 = time.perf_counter()
        e2e_run(e2e_steps)
        torch.cuda.synchronize()
        reps.append((time.perf_counter() - t0) * 1e3)
    e2e_ms = sorted(reps)[len(reps) // 2]
    e2e_value = shard.aggregate_throughput(cfg.B * e2e_steps, e2e_ms, dev)
    e2e = {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
           "d2h_bytes_per_step": pipe.d2h_bytes, "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
           "note": "pipeline.HostPipeline: pinned host calibration+depth+feat in (one packed H2D), "
                   "d_depth+d_feat out (one packed D2H), both copies inside the slot's CUDA graph, every step's "
                   "result read on the host, six steps in flight (copies overlap kernels); upstream dBEV "
                   "stays on the device; host wall clock, median of %d runs" % len(reps),
           "host_link_gbs": {"h2d_concurrent_per_rank": [round(x, 1) for x in link_all],
                             "h2d_concurrent_sum": round(sum(link_all), 1),
                             "h2d_plus_d2h_concurrent_per_rank": [round(x, 1) for x in bidir_all],
                             "h2d_plus_d2h_concurrent_sum": round(sum(bidir_all), 1)},
           "host_traffic_gbs": round(e2e_value / cfg.B * (pipe.h2d_bytes + pipe.d2h_bytes) / 1e9, 1),
           "host": host_info, "repeats_ms_per_step": [round(r / e2e_steps, 4) for r in reps]}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(cfg, world, args.sets),
        "run": {"in_flight": "%d independent batches in flight per GPU (one CUDA stream each)" % in_flight,
                "bev_layout": "channels_last (NHWC storage of the logical (B,C*Z,X,Y) map)",
                "launch": "stream launches" if args.no_graph else "cuda-graph replay (feature staging || plan)",
                "step": "camera prep+geometry+sort+intervals, feature staging, fused fwd, fused bwd"},
        "repeats": repeats,
        "block_ms": {"median": med_ms, "p10": all_ms[len(all_ms) // 10], "p90": all_ms[(9 * len(all_ms)) // 10]},
        "serial": {"ms_per_step": serial_ms, "value": cfg.B * world / (serial_ms * 1e-3), "unit": UNIT,
                   "note": "one batch in flight: every kernel of a step waits for the previous step"},
        "cached_plan": {"ms_per_step": cached_ms, "value": cfg.B * world / (cached_ms * 1e-3), "unit": UNIT,
                        "note": "one batch in flight, plan reused (fixed rig): feature staging + fwd + bwd"},
        "clocks": clocks.summary(), "e2e": e2e, "roofline": roofline, "roofline_bwd": roofline_bwd,
        "gpu_launches": KERNELS_PER_STEP * args.steps * repeats,
    }
    if module_api is not None:
        line["module_api"] = module_api
    if reference_gpu is not None:
        line["reference_gpu"] = reference_gpu
    if all_cpus:
        os.sched_setaffinity(0, all_cpus)     # the CPU baseline below uses every host core
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        r = cpu_port_run(cfg, steps=60, warmup=1, budget_s=12.0)
        line["cpu_baseline"] = {"value": r["samples_per_s"], "unit": UNIT, "cores": r["cores"], "kind": "port",
                                "sample": "%d steps of one %d-sample batch (C restatement of the reference "
                                          "algorithm, oracle/lss_oracle.c, OpenMP)" % (r["steps"], cfg.B)}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def reference_legs(cfg, dev, step0):
    """(module_api, reference_gpu): the reference's own LSS class (oracle/_ref via oracle/ref_import.py, image
    encoder replaced by Identity: the stage is fed the encoder's output shape (B*N, 512, fH, fW)), timed with the
    drop-in installed and as written.  Checker / baseline legs: nothing here is part of the product path."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import ref_import
        if not ref_import.available():
            return None, {"unavailable": "reference tree not staged (oracle/stage_ref.py)"}
        from lss2_multimodal_nu_b200 import patch, synthetic as S
    except Exception as e:  # noqa: BLE001
        return None, {"unavailable": str(e)[:120]}
    if cfg.C != 64:
        return None, {"unavailable": "the reference's LSS class hard-codes camC = 64 (src/model_baseline.py:25)"}
    cal = [torch.from_numpy(v).to(dev) for v in S.make_calibration(cfg, 1234).values()]
    gen = torch.Generator(device=dev); gen.manual_seed(99)
    x = torch.randn(cfg.B * cfg.N, 512, cfg.fH, cfg.fW, device=dev, generator=gen)
    dbev = step0.dbev.detach()

    def build():
        torch.manual_seed(0)
        return ref_import.build_lss(cfg.B, cfg.grid_conf(), cfg.data_aug_conf()).to(dev).train()

    def time_steps(model, n, warm):
        xg = x.clone().requires_grad_(True)
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)

        def one():
            xg.grad = None
            bev = model.get_voxels(xg, *cal)
            bev.backward(dbev)
        for _ in range(warm):
            one()
        torch.cuda.synchronize(dev)
        e0.record()
        for _ in range(n):
            one()
        e1.record()
        torch.cuda.synchronize(dev)
        return e0.elapsed_time(e1) / n

    with torch.cuda.device(dev):
        m = build()
        ref_ms = time_steps(m, 3, 1)
        patch.install(m)
        api_ms = time_steps(m, 200, 20)
        hits = patch._cache(m).prefetch_hits
    reference_gpu = {"value": cfg.B / (ref_ms * 1e-3), "unit": UNIT, "ms_per_step": ref_ms, "steps": 3,
                     "note": "UNMODIFIED reference LSS.get_voxels (get_geometry + CamEncode + voxel_pooling with "
                             "QuickCumsum, src/model_baseline.py:128-133) forward + backward, stock PyTorch on this GPU"}
    module_api = {"value": cfg.B / (api_ms * 1e-3), "unit": UNIT, "ms_per_step": api_ms, "steps": 200,
                  "speedup_vs_reference_gpu": ref_ms / api_ms,
                  "note": "the same LSS.get_voxels + backward with patch.install(model): eager, one stream, "
                          "per-call allocations, the reference's own 1x1 depthnet conv included; plan built inside "
                          "the call (no model-level forward here, so no prefetch: %d hits)" % hits}
    return module_api, reference_gpu


def main():
    args = parse()
    from lss2_multimodal_nu_b200 import synthetic as S
    cfg = S.config(args.config)
    if args.impl == "reference":
        run_reference_arm(args, cfg)
    else:
        run_ours(args, cfg)


if __name__ == "__main__":
    main()

#!/usr/bin/env python3
"""bench.py — Msamples/s of the path-tracing hot path on the Dragon scene (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload NAME]

A "step" is one full render of the workload frame (width*height*spp camera samples through the whole
wavefront pipeline).  N>1 is launched by torchrun, one rank per GPU: the frame is tile-partitioned over the
ranks, each rank renders its tiles, ONE NCCL all-gather per step moves the finished tile buffers and rank 0
scatters them into the frame (strong scaling: total work is fixed).

Our arm prints one JSON line:
  value     whole-job Msamples/s with the frame staying in HBM (CUDA events, max over ranks)
  e2e       same metric through the reference-facing call Renderer.Draw(host frame) (b200pt_render):
            options go in, the finished frame comes back to host memory, all inside the timed region
  roofline  dominant kernel vs the HBM roofline with the algorithmic-bytes formula of SURVEY.md §8d
  cpu_baseline  the reference's own CPU renderer (oracle/_ref) timed on this box on a bounded sample

--impl reference times the UNMODIFIED reference CPU renderer (csrt::Renderer::Draw built from
/root/reference into oracle/_ref, all host threads) on the same workload, each step a bounded spp sample.
"""
import argparse
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))

# name -> (pack, width, height, spp, description)
WORKLOADS = {
    "dragon-1024-256spp": ("dragon", 1024, 1024, 256, "resources/scene/dragon/scene.xml 1024x1024 256spp (BASELINE configs[1])"),
    "dragon-1080p-4096spp": ("dragon", 1920, 1080, 4096, "resources/scene/dragon/scene.xml 1920x1080 4096spp (BASELINE configs[4])"),
    "matpreview-1024-512spp": ("matpreview", 1024, 1024, 512, "resources/scene/matpreview/rough_conductor.xml 1024x1024 512spp (configs[2])"),
    "volumetric-1024-2048spp": ("volumetric-caustic", 1024, 1024, 2048, "resources/scene/volumetric-caustic/scene_v0.6.xml 1024x1024 2048spp (configs[3])"),
    "cornell-256-64spp": ("cornell-box", 256, 256, 64, "resources/scene/cornell-box/scene_v0.6.xml 256x256 64spp (smoke)"),
}
METRIC = "Msamples/sec at 1024^2 256spp (Dragon), 1/2/4/8 B200; per-pixel rel-L2 vs --cpu ref"
# bytes a cache-less traversal must move per unit of work
BYTES_PER_NODE_VISIT, BYTES_PER_PRIM_TEST, BYTES_PER_CLOSEST_RAY = 32, 36, 52
REF_SPP_PER_STEP = 8  # bounded sample of the reference arm / cpu_baseline


def pack_path(name):
    return os.path.join(ROOT, "scenes", name + ".b200scene")


class ClockSampler(threading.Thread):
    """Samples SM clock + throttle reasons with NVML while the timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.reasons, self.max_mhz, self.stop_flag = index, [], set(), None, False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.dev = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.dev, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self.stop_flag:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.dev, nv.NVML_CLOCK_SM))
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(self.dev)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.dev)
                for bit, name in names.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.05)

    def result(self):
        self.stop_flag = True
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": ["nvml unavailable"]}
        s = sorted(self.samples)
        return {"sm_mhz": s[len(s) // 2], "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def cpu_baseline(pack, width, height, spp_sample, spp_full):
    """The reference's CPU renderer on this box, bounded sample: `spp_sample` of the workload's spp."""
    import refcheck
    ref = refcheck.ref_lib("woop")
    r = refcheck.RefRenderer(ref, pack, width, height, spp_sample)
    seconds = r.draw()
    r.close()
    return {"value": width * height * spp_sample / seconds / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference",
            "sample": f"{width}x{height} at {spp_sample} of {spp_full} spp, one csrt::Renderer::Draw, {seconds:.2f} s "
                      f"(scene commit {r.build_seconds:.1f} s excluded)"}


def reference_gpu_baseline(pack, width, height, spp_sample, spp_full):
    """The reference's OWN CUDA backend (its one-thread-per-pixel megakernel DispathRaysCuda, renderer.cpp:88-95, rebuilt for
    sm_100a into oracle/_ref/libcsrt_ref_cuda.so) on this GPU, bounded sample.  A comparator like cpu_baseline, not a target."""
    import ctypes
    import numpy as np
    import __graft_entry__ as ge
    path = os.path.join(ROOT, "oracle", "_ref", "libcsrt_ref_cuda.so")
    if not os.path.exists(path):
        return {"value": None, "unit": "Msamples/s", "sample": "unavailable: oracle/_ref/libcsrt_ref_cuda.so not built"}
    L = ctypes.CDLL(path)
    L.ref_create_cuda.restype = ctypes.c_void_p
    L.ref_create_cuda.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(ctypes.c_double)]
    L.ref_draw_cuda.restype = ctypes.c_double
    L.ref_draw_cuda.argtypes = [ctypes.c_void_p, ctypes.c_void_p]
    L.ref_destroy_cuda.argtypes = [ctypes.c_void_p]
    L.ref_last_error.restype = ctypes.c_char_p
    scene = ge.load_package().Scene(pack)
    build = ctypes.c_double()
    handle = L.ref_create_cuda(scene.desc, width, height, spp_sample, ctypes.byref(build))
    if not handle:
        return {"value": None, "unit": "Msamples/s", "sample": "unavailable: " + L.ref_last_error().decode(errors="replace")}
    frame = np.zeros((height, width, 3), dtype=np.float32)
    times = [L.ref_draw_cuda(handle, frame.ctypes.data) for _ in range(2)]  # the first Draw pages the managed scene in
    L.ref_destroy_cuda(handle)
    if min(times) <= 0:
        return {"value": None, "unit": "Msamples/s", "sample": "unavailable: " + L.ref_last_error().decode(errors="replace")}
    return {"value": width * height * spp_sample / min(times) / 1e6, "unit": "Msamples/s", "kind": "reference --gpu backend, sm_100a build",
            "sample": f"{width}x{height} at {spp_sample} of {spp_full} spp, csrt::Renderer::Draw on BackendType::kCuda, best of 2: {min(times):.3f} s"}


def run_reference(args, workload):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, width, height, spp, desc = workload
    import refcheck
    try:
        ref = refcheck.ref_lib("woop")
    except FileNotFoundError as e:
        print(json.dumps({"impl": "reference", "unavailable": f"oracle/_ref not built: {e}"}))
        return
    r = refcheck.RefRenderer(ref, pack_path(name), width, height, REF_SPP_PER_STEP)
    for _ in range(args.warmup):
        r.draw()
    t = [r.draw() for _ in range(args.steps)]
    r.close()
    samples = width * height * REF_SPP_PER_STEP
    value = samples * args.steps / sum(t) / 1e6
    sample = (f"each step = {width}x{height} at {REF_SPP_PER_STEP} of {spp} spp through csrt::Renderer::Draw "
              f"(renderer.cpp:678), {os.cpu_count()} host threads; scene commit {r.build_seconds:.1f} s excluded")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "reference scene (Dragon), parsed by the reference's XML parser",
        "config": {"workload": desc, "sample": sample},
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": sample},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def run_b200(args, workload):
    import numpy as np
    import torch
    import torch.distributed as dist
    import __graft_entry__ as ge

    pkg = ge.load_package()
    name, width, height, spp, desc = workload
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; the b200 arm has no CPU fallback")
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    scene = pkg.Scene(pack_path(name))
    t0 = time.time()
    renderer = pkg.Renderer(scene, device=local_rank)
    create_s = time.time() - t0
    stream = torch.cuda.current_stream()
    frame = torch.zeros(height * width * 3, dtype=torch.float32, device="cuda")
    tile_floats = pkg.tile_buffer_floats(width, height, world)
    tiles = torch.zeros(tile_floats, dtype=torch.float32, device="cuda")
    gathered = torch.zeros(tile_floats * world, dtype=torch.float32, device="cuda") if world > 1 else None
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2

    def step(stats=0):
        """One full frame, result left in HBM (on rank 0 for N>1)."""
        if world == 1:
            renderer.draw_device(frame, width, height, spp, seed=1, stream=stream.cuda_stream, stats=stats)
        else:
            renderer.draw_tiles_device(tiles, rank, world, width, height, spp, seed=1, stream=stream.cuda_stream, stats=stats)
            dist.all_gather_into_tensor(gathered, tiles)
            if rank == 0:
                renderer.assemble_tiles_device(gathered, frame, width, height, world, stream=stream.cuda_stream)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        step()
    barrier()

    # ---- timed region: K steps, CUDA events per step on the launching stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    step_ms = []
    for _ in range(args.steps):
        flush.zero_()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step()
        e1.record(stream)
        barrier()
        step_ms.append(e0.elapsed_time(e1))
    clocks = sampler.result()
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    total_ms = float(total_ms.item())
    samples_per_step = width * height * spp
    value = samples_per_step * args.steps / total_ms / 1e3
    launches = renderer.stats()["kernel_launches"] * args.steps

    # ---- e2e: the call a user of the reference makes — Draw(host frame) — every step ----
    e2e = None
    if world == 1:
        host_frame = np.zeros((height, width, 3), dtype=np.float32)
        renderer.Draw(host_frame, width, height, spp, seed=1)
        t_e2e = []
        for _ in range(args.steps):
            flush.zero_()
            torch.cuda.synchronize()
            t = time.perf_counter()
            renderer.Draw(host_frame, width, height, spp, seed=1)  # blocking; frame is in host memory on return
            t_e2e.append(time.perf_counter() - t)
        e2e = {"value": samples_per_step * args.steps / sum(t_e2e) / 1e6, "unit": "Msamples/s",
               "h2d_bytes_per_step": 40, "d2h_bytes_per_step": host_frame.nbytes,
               "note": "b200pt_render(): render options in (40 B), finished frame copied back to the caller's host buffer; "
                       "the scene is resident in HBM from b200pt_create (as the reference keeps it from Renderer())"}
    else:
        # N>1: device render + NCCL gather + D2H of the assembled frame on rank 0, wall clock max over ranks
        host_pinned = torch.empty(height * width * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
        t_e2e = []
        for _ in range(args.steps):
            barrier()
            t = time.perf_counter()
            step()
            if rank == 0:
                host_pinned.copy_(frame, non_blocking=False)
            barrier()
            t_e2e.append(time.perf_counter() - t)
        tt = torch.tensor([sum(t_e2e)], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        e2e = {"value": samples_per_step * args.steps / float(tt.item()) / 1e6, "unit": "Msamples/s", "h2d_bytes_per_step": 40,
               "d2h_bytes_per_step": height * width * 12}

    # ---- the same frame with the tile visibility pre-pass off (every sample traces a camera ray), for transparency ----
    no_cull_ms = []
    if world == 1:
        for _ in range(3):
            flush.zero_()
            torch.cuda.synchronize()
            renderer.draw_device(frame, width, height, spp, seed=1, stream=stream.cuda_stream, flags=pkg.RENDER_NO_TILE_CULL)
            torch.cuda.synchronize()
            no_cull_ms.append(renderer.stats()["render_ms"])

    # ---- roofline of the dominant kernel: one counted + one event-timed step (not part of `value`) ----
    step(stats=pkg.STATS_COUNTERS)
    barrier()
    counted = renderer.stats()
    step(stats=pkg.STATS_TIMING)
    barrier()
    timed = renderer.stats()
    def algorithmic_bytes(c, closest):
        return (BYTES_PER_NODE_VISIT * c["node_visits"] + BYTES_PER_PRIM_TEST * c["prim_tests"]
                + (BYTES_PER_CLOSEST_RAY * c["rays"] if closest else 0))

    def gbps(nbytes, ms):
        return nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0

    # k_primary traces the camera rays; k_trace traces, in ONE launch per bounce, the bounce rays (closest hit, counted
    # under "extend") and the NEE rays (any hit, counted under "shadow"); its time is reported under "extend".
    alg_primary = algorithmic_bytes(counted["primary"], True)
    alg_trace = algorithmic_bytes(counted["extend"], True) + algorithmic_bytes(counted["shadow"], False)
    kernels = {
        "primary": {"ms": timed["primary"]["ms"], "launches": timed["primary"]["launches"], "rays": counted["primary"]["rays"],
                    "algorithmic_bytes": alg_primary, "GBps": gbps(alg_primary, timed["primary"]["ms"])},
        "trace": {"ms": timed["extend"]["ms"], "launches": timed["extend"]["launches"], "closest_hit_rays": counted["extend"]["rays"],
                  "any_hit_rays": counted["shadow"]["rays"], "algorithmic_bytes": alg_trace, "GBps": gbps(alg_trace, timed["extend"]["ms"])},
        "shade": {"ms": timed["shade"]["ms"], "launches": timed["shade"]["launches"]},
        "other": {"ms": timed["other"]["ms"], "launches": timed["other"]["launches"]},
        "tail": {"ms": timed["tail"]["ms"], "launches": timed["tail"]["launches"]},
    }
    dominant = max(("primary", "trace"), key=lambda k: kernels[k]["ms"])
    peak, peak_src = measured_peak()
    dk = kernels[dominant]
    roofline = {"bound": "hbm", "kernel": "k_" + dominant, "achieved": dk["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dk["GBps"] / peak, "traffic": None, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": dk["algorithmic_bytes"] / max(1, dk["launches"]),
                "avg_launch_ms": dk["ms"] / max(1, dk["launches"]), "share_of_step": dk["ms"] / timed["render_ms"],
                "kernels": kernels,
                "formula": "32 B x child-box tests + 36 B x primitive tests + 52 B x closest-hit rays (SURVEY.md §8d), counted by the kernels themselves"}
    # DRAM bytes of the dominant kernel from the committed ncu capture of the same frame (tools/gpu_evidence.sh), per launch of
    # THIS step: the capture's per-frame total over this step's launch count (the capture may run more, smaller launches).
    traffic_file = os.path.join(ROOT, "profiles", "r01_dram_traffic.json")
    if os.path.exists(traffic_file) and name == "dragon" and (width, height, spp) == (1024, 1024, 256):
        with open(traffic_file) as f:
            detail = json.load(f).get("k_" + dominant + "_detail")
        if detail:
            roofline["traffic"] = (detail["dram_read_bytes_total"] + detail["dram_write_bytes_total"]) / max(1, dk["launches"])
            roofline["traffic_source"] = "profiles/r01_dram_traffic.json (ncu dram__bytes_read.sum + dram__bytes_write.sum over one frame)"

    if rank == 0:
        cpu = None
        if not args.no_cpu_baseline and world == 1:  # the contract asks for it at N=1 only
            try:
                cpu = cpu_baseline(pack_path(name), width, height, REF_SPP_PER_STEP * 2, spp)
            except Exception as e:  # the reference build is test infrastructure; its absence must not hide the GPU number
                cpu = {"value": None, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": f"unavailable: {e}"}
        ref_gpu = None
        if not args.no_cpu_baseline and world == 1:
            # in a child process: the comparator is foreign code on the same GPU and must not be able to take this line down
            renderer.close()
            try:
                import subprocess
                probe = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--ref-gpu-probe"],
                                       capture_output=True, text=True, timeout=600)
                ref_gpu = json.loads(probe.stdout.strip().splitlines()[-1])
            except Exception as e:
                ref_gpu = {"value": None, "unit": "Msamples/s", "sample": f"unavailable: {e}"}
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference scene (Dragon, 831812 triangles) parsed once by the reference's XML parser into scenes/dragon.b200scene",
            "config": {"workload": desc, "tile_split": f"{world} rank(s), 8x8-pixel tiles dealt round-robin, one NCCL all-gather per step" if world > 1 else "single GPU",
                       "l2": "256 MB buffer written between timed steps (L2 flush); scene (160 MB) + wavefront state (4 GB) also exceed the 126 MB L2",
                       "rng": "Philox4x32-10 keyed by seed, counter = (pixel, sample, depth)", "scene_create_s": create_s,
                       "tile_visibility_prepass": {
                           "what": "8x8 screen tiles whose camera-ray pyramid provably misses every box of a 384-box BVH cut are not traced "
                                   "(exact: the frame is bit-identical, tests/test_gpu_parity.py); the pre-pass runs inside the timed region",
                           "active_tiles": counted["active_tiles"], "local_tiles": counted["local_tiles"],
                           "value_with_prepass_off": (samples_per_step / min(no_cull_ms) / 1e3) if no_cull_ms else None}},
            "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu,
            "reference_gpu_baseline": ref_gpu,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    renderer.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dragon-1024-256spp", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--ref-gpu-probe", action="store_true", help=argparse.SUPPRESS)  # child process of the b200 arm
    args = ap.parse_args()
    workload = WORKLOADS[args.workload]
    if not os.path.exists(pack_path(workload[0])):
        raise SystemExit(f"scene pack {pack_path(workload[0])} missing; run __graft_entry__.build() where /root/reference exists")
    if args.ref_gpu_probe:
        name, width, height, spp, _ = workload
        print(json.dumps(reference_gpu_baseline(pack_path(name), width, height, REF_SPP_PER_STEP * 2, spp)))
    elif args.impl == "reference":
        run_reference(args, workload)
    else:
        run_b200(args, workload)


if __name__ == "__main__":
    main()

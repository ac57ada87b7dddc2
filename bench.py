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
    "mercury-256-32spp": ("mercury", 256, 256, 32, "resources/scene/mercury/smooth_diffuse.xml 256x256 32spp (BASELINE configs[0])"),
    "cornell-256-64spp": ("cornell-box", 256, 256, 64, "resources/scene/cornell-box/scene_v0.6.xml 256x256 64spp (smoke)"),
}
# (label, workload, timed steps, golden) measured after the headline: N=1 runs the single-GPU configs and C5, N>1 runs C5
OTHER_CONFIGS_N1 = [("C1", "mercury-256-32spp", 5, "fullsize_mercury.npz"), ("C3", "matpreview-1024-512spp", 3, "fullsize_matpreview.npz"),
                    ("C4", "volumetric-1024-2048spp", 2, "fullsize_volumetric-caustic.npz"), ("C5", "dragon-1080p-4096spp", 2, "fullsize_dragon-1080p.npz")]
OTHER_CONFIGS_MULTI = [("C5", "dragon-1080p-4096spp", 3, "fullsize_dragon-1080p.npz")]
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
    """The reference's CPU renderer on this box, bounded sample: `spp_sample` of the workload's spp.  Both builds of the same
    unmodified sources: the stock one (CMake Release flags) is `value`, the x86-64-v3 unity build ("the fairer
    comparator") is `fast_build`."""
    import refcheck
    out = None
    for variant in ("woop", "fast"):
        if variant == "fast" and not fast_build_usable():
            continue
        r = refcheck.RefRenderer(refcheck.ref_lib(variant), pack, width, height, spp_sample)
        seconds = r.draw()
        r.close()
        rec = {"value": width * height * spp_sample / seconds / 1e6, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference",
               "sample": f"{width}x{height} at {spp_sample} of {spp_full} spp, one csrt::Renderer::Draw, {seconds:.2f} s "
                         f"(scene commit {r.build_seconds:.1f} s excluded)"}
        if variant == "woop":
            rec["build"] = "g++ -O3 -DNDEBUG (CMake Release), -DWATERTIGHT_TRIANGLES"
            out = rec
        else:
            out["fast_build"] = {"value": rec["value"], "unit": "Msamples/s", "sample": rec["sample"],
                                 "build": "-O3 -march=x86-64-v3, hot-path sources as one translation unit (oracle/Makefile FAST_OPT)"}
    return out


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


def host_threads_note():
    return f"{os.cpu_count()} host threads"


def fast_build_usable():
    """libcsrt_ref_fast.so is compiled for x86-64-v3 (AVX2 + FMA): only load it on a host that has them."""
    try:
        flags = open("/proc/cpuinfo").read()
    except OSError:
        return False
    return all(f" {x}" in flags for x in ("avx2", "fma", "bmi2")) and os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libcsrt_ref_fast.so"))


def shared_config(desc, width, height, spp):
    """The part of `config` both arms print identically: the workload both measure."""
    return {"workload": desc, "width": width, "height": height, "spp": spp,
            "l2": "b200 arm: 256 MB buffer written between timed steps (L2 flush), and scene (160 MB) + wavefront state (~190 B per sample slot in flight: GBs) exceed the 126 MB L2; reference arm: CPU"}


def run_reference(args, workload):
    """The reference's own CPU renderer (unmodified sources, oracle/_ref), all host threads, bounded spp sample per step.
    `value` comes from the stock build (CMake Release flags, -O3); the x86-64-v3 unity build of the same sources (BASELINE.md §3.1:
    "the fairer comparator") is timed beside it and reported in cpu_baseline.fast_build."""
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
    fast = None
    if fast_build_usable():
        try:
            rf = refcheck.RefRenderer(refcheck.ref_lib("fast"), pack_path(name), width, height, REF_SPP_PER_STEP)
            rf.draw()
            tf = [rf.draw() for _ in range(min(3, args.steps))]
            rf.close()
            fast = {"value": samples * len(tf) / sum(tf) / 1e6, "unit": "Msamples/s", "build": "-O3 -march=x86-64-v3, hot-path sources as one translation unit (oracle/Makefile FAST_OPT)"}
        except Exception as e:
            fast = {"value": None, "note": f"unavailable: {e}"}
    sample = (f"each step = {width}x{height} at {REF_SPP_PER_STEP} of {spp} spp through csrt::Renderer::Draw "
              f"(renderer.cpp:678), {host_threads_note()}; scene commit {r.build_seconds:.1f} s excluded; throughput is linear in spp")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * sum(t) / args.steps, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "reference scene (Dragon), parsed by the reference's XML parser",
        "config": shared_config(desc, width, height, spp),
        "cpu_baseline": {"value": value, "unit": "Msamples/s", "cores": os.cpu_count(), "kind": "reference", "sample": sample,
                         "build": "g++ -O3 -DNDEBUG (CMake Release), -DWATERTIGHT_TRIANGLES", "fast_build": fast},
        "e2e": {"value": value, "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }))


def rel_l2(a, b):
    import numpy as np
    return float(np.linalg.norm(a.astype(np.float64) - b) / np.linalg.norm(b.astype(np.float64)))


def boxed(f, box=8):
    h, w = f.shape[:2]
    return f[: h // box * box, : w // box * box].reshape(h // box, box, w // box, box, 3).mean(axis=(1, 3))


class Job:
    """One workload on this rank: renderer + the buffers of a step (a frame for N=1; tiles + gather + assemble for N>1)."""

    def __init__(self, pkg, pack, width, height, world, rank, local_rank):
        import torch
        self.pkg, self.torch, self.width, self.height, self.world, self.rank = pkg, torch, width, height, world, rank
        self.scene = pkg.Scene(pack)
        t0 = time.time()
        self.renderer = pkg.Renderer(self.scene, device=local_rank)
        self.create_s = time.time() - t0
        self.stream = torch.cuda.current_stream()
        self.frame = torch.zeros(height * width * 3, dtype=torch.float32, device="cuda")
        n = pkg.tile_buffer_floats(width, height, world)
        self.tiles = torch.zeros(n, dtype=torch.float32, device="cuda") if world > 1 else None
        self.gathered = torch.zeros(n * world, dtype=torch.float32, device="cuda") if world > 1 else None

    def step(self, spp, seed=1, stats=0, flags=0):
        """One full frame, result left in HBM (on rank 0 for N>1)."""
        r, s = self.renderer, self.stream.cuda_stream
        if self.world == 1:
            r.draw_device(self.frame, self.width, self.height, spp, seed=seed, stream=s, stats=stats, flags=flags)
        else:
            import torch.distributed as dist
            r.draw_tiles_device(self.tiles, self.rank, self.world, self.width, self.height, spp, seed=seed, stream=s, stats=stats)
            dist.all_gather_into_tensor(self.gathered, self.tiles)
            if self.rank == 0:
                r.assemble_tiles_device(self.gathered, self.frame, self.width, self.height, self.world, stream=s)

    def barrier(self):
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.torch.cuda.synchronize()

    def timed_steps(self, spp, steps, flush):
        """CUDA events on the launching stream around each step, barrier + synchronize on both sides, L2 flushed in between;
        returns the total over the steps, max over ranks."""
        torch = self.torch
        ms = []
        for _ in range(steps):
            flush.zero_()
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(self.stream)
            self.step(spp)
            e1.record(self.stream)
            self.barrier()
            ms.append(e0.elapsed_time(e1))
        total = torch.tensor([sum(ms)], dtype=torch.float64, device="cuda")
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(total, op=dist.ReduceOp.MAX)
        return float(total.item())

    def frame_host(self):
        return self.frame.cpu().numpy().reshape(self.height, self.width, 3)

    def parity(self, golden_file):
        """Frame of this job vs a frame of the UNMODIFIED reference CPU renderer at the same size and spp (tests/golden/fullsize_*.npz,
        made by tests/golden/make_fullsize_golden.py): whole-image mean ratio, per-pixel and 8x8-box rel-L2, and the same two between
        two of our own seeds (the Monte Carlo noise floor the comparison cannot go below)."""
        import numpy as np
        path = os.path.join(ROOT, "tests", "golden", golden_file)
        if not os.path.exists(path):
            return None
        g = np.load(path)
        golden = g["frame"].astype(np.float32)
        w, h, spp = (int(x) for x in g["size"])
        if (w, h) != (self.width, self.height):
            return None
        self.step(spp, seed=31)
        self.barrier()
        a = self.frame_host() if self.rank == 0 else None
        self.step(spp, seed=32)
        self.barrier()
        if self.rank != 0:
            return None
        b = self.frame_host()
        # Exact mode (test hook b200pt_debug_render_replay, tests/test_gpu_replay.py): the reference's per-pixel LCG stream and loop
        # shape around the product's device functions -> the SAME samples as the golden frame, compared per pixel.  The golden is
        # stored as float16 (5e-4 relative rounding), hence the 2e-3 tolerance.
        exact = None
        try:
            e = self.renderer.render_replay(w, h, spp)
            d = np.abs(e.astype(np.float64) - golden).max(axis=2) / np.maximum(np.abs(golden).max(axis=2), 1e-3)
            exact = {"what": "same LCG stream as the reference: per-pixel agreement, not statistics",
                     "pixels_within_2e-3": float(np.mean(d <= 2e-3)), "pixels_within_1e-2": float(np.mean(d <= 1e-2)),
                     "rel_l2": rel_l2(e, golden), "rel_l2_of_float16_rounding": rel_l2(e.astype(np.float16).astype(np.float32), e)}
        except self.pkg.MyException as err:
            exact = {"unavailable": str(err)}
        return {"against": f"tests/golden/{golden_file}: csrt::Renderer::Draw (reference CPU build), {w}x{h} at {spp} spp on both sides",
                "exact_mode": exact,
                "mean_ratio": float(a.mean() / golden.mean()), "rel_l2_pixel": rel_l2(a, golden), "rel_l2_box8": rel_l2(boxed(a), boxed(golden)),
                "noise_floor_pixel": rel_l2(a, b), "noise_floor_box8": rel_l2(boxed(a), boxed(b)),
                "tolerance": "mean within 0.5 %, box8 <= 2 x floor + 0.5 %, pixel <= 1.5 x floor + 0.5 % (tests/test_gpu_parity.py)"}

    def kernel_split(self, spp):
        """Per-kernel-class CUDA-event times of one extra step (B200PT_STATS_TIMING renders with ONE arena so that launches are
        serial), and the traversal counters of another."""
        self.step(spp, stats=self.pkg.STATS_COUNTERS)
        self.barrier()
        counted = self.renderer.stats()
        self.step(spp, stats=self.pkg.STATS_TIMING)
        self.barrier()
        return counted, self.renderer.stats()

    def close(self):
        self.renderer.close()
        self.frame = self.tiles = self.gathered = None
        self.torch.cuda.empty_cache()


def algorithmic_bytes(c, closest):
    return (BYTES_PER_NODE_VISIT * c["node_visits"] + BYTES_PER_PRIM_TEST * c["prim_tests"]
            + (BYTES_PER_CLOSEST_RAY * c["rays"] if closest else 0))


def gbps(nbytes, ms):
    return nbytes / (ms * 1e-3) / 1e9 if ms > 0 else 0.0


def kernel_table(counted, timed):
    # k_primary traces the camera rays; k_trace traces, in ONE launch per bounce, the bounce rays (closest hit, counted
    # under "extend") and the NEE rays (any hit, counted under "shadow"); its time is reported under "extend".
    alg_primary = algorithmic_bytes(counted["primary"], True)
    alg_trace = algorithmic_bytes(counted["extend"], True) + algorithmic_bytes(counted["shadow"], False)
    return {
        "primary": {"ms": timed["primary"]["ms"], "launches": timed["primary"]["launches"], "rays": counted["primary"]["rays"],
                    "box_tests": counted["primary"]["node_visits"], "primitive_tests": counted["primary"]["prim_tests"],
                    "algorithmic_bytes": alg_primary, "GBps": gbps(alg_primary, timed["primary"]["ms"])},
        "trace": {"ms": timed["extend"]["ms"], "launches": timed["extend"]["launches"], "closest_hit_rays": counted["extend"]["rays"],
                  "any_hit_rays": counted["shadow"]["rays"],
                  "box_tests": {"closest_hit": counted["extend"]["node_visits"], "any_hit": counted["shadow"]["node_visits"]},
                  "primitive_tests": {"closest_hit": counted["extend"]["prim_tests"], "any_hit": counted["shadow"]["prim_tests"]},
                  "algorithmic_bytes": alg_trace, "GBps": gbps(alg_trace, timed["extend"]["ms"])},
        "shade": {"ms": timed["shade"]["ms"], "launches": timed["shade"]["launches"]},
        "other": {"ms": timed["other"]["ms"], "launches": timed["other"]["launches"]},
        "tail": {"ms": timed["tail"]["ms"], "launches": timed["tail"]["launches"]},
    }


def ncu_counters(pack, width, height, spp, kernel_prefix):
    """What ncu measured for this kernel on this workload (committed capture, tools/ncu_counters.py): the bytes that really
    moved and how busy the issue slots were — the numbers that say what binds, next to the contract's byte formula."""
    import glob
    candidates = [os.path.join(ROOT, "profiles", f"r02_counters_{pack}_{width}x{height}x{spp_try}.json") for spp_try in (spp, 256, 128, 64)]
    # no capture at this frame size: one of the same scene at another size still says what binds the kernel (rates and ratios,
    # not bytes per launch)
    candidates += sorted(glob.glob(os.path.join(ROOT, "profiles", f"r02_counters_{pack}_*.json")))
    for path in candidates:
        if os.path.exists(path):
            with open(path) as f:
                doc = json.load(f)
            rows = {k: v for k, v in doc["kernels"].items() if k.startswith(kernel_prefix)}
            if not rows:
                return None
            k, v = max(rows.items(), key=lambda kv: kv[1]["time_ms"])
            try:
                spp_capture = int(os.path.basename(path)[:-len(".json")].rsplit("x", 1)[1])
            except (IndexError, ValueError):
                continue  # not a <scene>_<w>x<h>x<spp>.json capture
            return {"file": os.path.relpath(path, ROOT), "kernel": k, "capture": doc.get("_what", ""), "spp": spp_capture,
                    "same_size": f"_{width}x{height}x" in os.path.basename(path), **v}
    return None


def roofline_block(pack, width, height, spp, kernels, render_ms):
    dominant = max(("primary", "trace"), key=lambda k: kernels[k]["ms"])
    peak, peak_src = measured_peak()
    dk = kernels[dominant]
    block = {"bound": "hbm", "kernel": "k_" + dominant, "achieved": dk["GBps"], "peak": peak, "unit": "GB/s",
             "frac": dk["GBps"] / peak, "traffic": None, "peak_source": peak_src,
             "algorithmic_bytes_per_launch": dk["algorithmic_bytes"] / max(1, dk["launches"]),
             "avg_launch_ms": dk["ms"] / max(1, dk["launches"]), "share_of_step": dk["ms"] / render_ms,
             "formula": "32 B x child-box tests + 36 B x primitive tests + 52 B x closest-hit rays (SURVEY.md §8d), counted by the kernels themselves"}
    c = ncu_counters(pack, width, height, spp, "k_" + dominant)
    if c:
        dram = c["dram_read_bytes"] + c["dram_write_bytes"]
        if c["same_size"]:
            # per launch LIKE `achieved`: the captured frame's DRAM bytes, scaled to this step's spp (traffic is linear in spp),
            # over this step's launch count (the timed split renders with one arena; the capture may have used more launches)
            block["traffic"] = dram * (spp / c["spp"]) / max(1, dk["launches"])
        block.update({"dram_gbs": c["dram_gbs"], "dram_frac": c["dram_gbs"] / peak, "l2_gbs": c["l2_gbs"], "issue_active": c["issue_active"],
                      "active_lanes": c["active_lanes"], "warps_active": c["warps_active"], "l1_hit": c["l1_hit"], "l2_hit": c["l2_hit"],
                      "counters_source": f"{c['file']} ({c['kernel']}, {c['launches']} launches, {c['time_ms']:.2f} ms under ncu): "
                                         + ("DRAM bytes of the captured frame x (spp of this step / spp of the capture) / this step's launch count = traffic" if c["same_size"] else
                                            "same scene and kernel at another frame size: rates and ratios only, no bytes per launch")})
        # what the counters say binds: HBM only when the DRAM pipe is actually busy
        if block["dram_frac"] < 0.5:
            block["bound"] = "issue/latency"
            block["bound_note"] = (f"DRAM moves {c['dram_gbs']:.0f} GB/s = {100 * block['dram_frac']:.1f} % of the measured peak while the §8d formula "
                                   f"counts {dk['GBps']:.0f} GB/s of node and triangle fetches: they are served by L1 ({100 * c['l1_hit']:.0f} % hit) and L2; the kernel "
                                   f"issues on {100 * c['issue_active']:.0f} % of the cycles with {c['active_lanes']:.1f} of 32 lanes active")
    else:
        block["bound_note"] = "no ncu capture of this scene under profiles/: the fraction is the §8d formula alone and does not say what binds"
    return block


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
    # scene_create_s times b200pt_create of the workload's scene.  What a process pays ONCE, whichever scene comes first, is timed
    # apart: its CUDA context (torch.zeros below) and the first b200pt_create (loading the library's kernels, streams, events,
    # pinned buffers: `library_init_s`, measured on the smallest scene).
    torch.zeros(1, device="cuda")
    torch.cuda.synchronize()
    t0 = time.time()
    first = pkg.Renderer(pkg.Scene(pack_path("cornell-box")), device=local_rank)
    first.close()
    library_init_s = time.time() - t0

    job = Job(pkg, pack_path(name), width, height, world, rank, local_rank)
    renderer, stream, frame = job.renderer, job.stream, job.frame
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")  # > 126 MB L2
    barrier = job.barrier

    warmup = max(args.warmup, 3)
    for _ in range(warmup):
        job.step(spp)
    barrier()

    # ---- timed region: K steps, CUDA events per step on the launching stream, L2 flushed between steps ----
    sampler = ClockSampler(local_rank)
    sampler.start()
    total_ms = job.timed_steps(spp, args.steps, flush)
    clocks = sampler.result()
    samples_per_step = width * height * spp
    value = samples_per_step * args.steps / total_ms / 1e3
    launches = renderer.stats()["kernel_launches"] * args.steps

    # ---- e2e: the call a user of the reference makes — Draw(host frame) — every step ----
    e2e = None
    if world == 1:
        # the caller's frame buffer is page-locked host memory (what a host that wants its frames fast allocates): the copy out of
        # b200pt_render is then one DMA, not a staged copy through the driver's bounce buffer
        host_frame = torch.zeros(height * width * 3, dtype=torch.float32).pin_memory().numpy().reshape(height, width, 3)
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
               "note": "b200pt_render(): render options in (40 B), finished frame copied back to the caller's (pinned) host buffer; "
                       "the scene is resident in HBM from b200pt_create (as the reference keeps it from Renderer())"}
    else:
        # N>1: device render + NCCL gather + D2H of the assembled frame on rank 0, wall clock max over ranks
        host_pinned = torch.empty(height * width * 3, dtype=torch.float32).pin_memory() if rank == 0 else None
        t_e2e = []
        for _ in range(args.steps):
            barrier()
            t = time.perf_counter()
            job.step(spp)
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
    counted, timed = job.kernel_split(spp)
    kernels = kernel_table(counted, timed)
    roofline = roofline_block(name, width, height, spp, kernels, timed["render_ms"])
    roofline["kernels"] = kernels
    parity = job.parity("fullsize_dragon.npz") if (name, width, height) == ("dragon", 1024, 1024) else None
    create_s = job.create_s
    job.close()

    # ---- the other BASELINE configs, same measurement in short form (not part of `value`) ----
    others = {}
    if not args.no_other_configs:
        plan = OTHER_CONFIGS_N1 if world == 1 else OTHER_CONFIGS_MULTI
        for label, key, steps, golden in plan:
            if key == args.workload:
                continue
            o_name, o_w, o_h, o_spp, o_desc = WORKLOADS[key]
            if not os.path.exists(pack_path(o_name)):
                others[label] = {"workload": o_desc, "unavailable": f"scenes/{o_name}.b200scene missing"}
                continue
            try:
                oj = Job(pkg, pack_path(o_name), o_w, o_h, world, rank, local_rank)
                oj.step(o_spp)
                oj.barrier()
                ms = oj.timed_steps(o_spp, steps, flush) / steps
                o_counted, o_timed = oj.kernel_split(o_spp)
                o_kernels = kernel_table(o_counted, o_timed)
                rec = {"workload": o_desc, "n_gpus": world, "Msamples_s": o_w * o_h * o_spp / ms / 1e3, "ms_per_step": ms, "steps": steps, "warmup": 1,
                       "scene_create_s": oj.create_s, "gpu_launches_per_step": oj.renderer.stats()["kernel_launches"],
                       "kernel_ms_single_arena": {k: round(v["ms"], 3) for k, v in o_kernels.items()},
                       "rays": {"primary": o_counted["primary"]["rays"], "closest_hit": o_counted["extend"]["rays"], "any_hit": o_counted["shadow"]["rays"]},
                       "roofline": roofline_block(o_name, o_w, o_h, o_spp, o_kernels, o_timed["render_ms"]),
                       "parity": oj.parity(golden) if golden else None}
                oj.close()
                others[label] = rec
            except Exception as e:  # a sub-record must not take the headline down
                others[label] = {"workload": o_desc, "error": f"{type(e).__name__}: {e}"}

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
            try:
                import subprocess
                probe = subprocess.run([sys.executable, os.path.abspath(__file__), "--workload", args.workload, "--ref-gpu-probe"],
                                       capture_output=True, text=True, timeout=600)
                ref_gpu = json.loads(probe.stdout.strip().splitlines()[-1])
            except Exception as e:
                ref_gpu = {"value": None, "unit": "Msamples/s", "sample": f"unavailable: {e}"}
        config = shared_config(desc, width, height, spp)
        config.update({"tile_split": f"{world} rank(s), 8x8-pixel tiles dealt round-robin, one NCCL all-gather per step" if world > 1 else "single GPU",
                       "rng": "Philox4x32-10 keyed by seed, counter = (pixel, sample, depth)", "scene_create_s": create_s, "library_init_s": library_init_s,
                       "tile_visibility_prepass": {
                           "what": "8x8 screen tiles whose camera-ray pyramid provably misses every box of a 384-box BVH cut, and pixels whose "
                                   "pyramid misses every box of a 4096-box cut, are not traced "
                                   "(exact: the frame is bit-identical, tests/test_gpu_parity.py); the pre-pass runs inside the timed region",
                           "active_tiles": counted["active_tiles"], "local_tiles": counted["local_tiles"], "active_pixels": counted["active_pixels"],
                           "value_with_prepass_off": (samples_per_step / min(no_cull_ms) / 1e3) if no_cull_ms else None}})
        line = {
            "metric": METRIC, "value": value, "unit": "Msamples/s", "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "reference scene (Dragon, 831812 triangles) parsed once by the reference's XML parser into scenes/dragon.b200scene",
            "config": config, "clocks": clocks, "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "parity": parity,
            "cpu_baseline": cpu, "reference_gpu_baseline": ref_gpu, "configs": others,
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="dragon-1024-256spp", choices=sorted(WORKLOADS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-other-configs", action="store_true", help="skip the C1 / C3 / C4 / C5 sub-records")
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

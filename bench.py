#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 draw path (BASELINE.json: Mtris/s and Gfragments/s for
vkCmdDrawIndexed at 4K; ms/frame at 1/2/4/8 B200).

Workload (config.workload = "C3/M1"): SURVEY §8(d) C3 — one vkCmdDrawIndexed of a 1000x500-quad grid
(1,000,000 triangles, 3,000,000 u32 indices, 501,501 vertices {vec4 pos, vec4 rgba}) at 3840x2160,
R8G8B8A8_UNORM + D32_SFLOAT, depth LESS_OR_EQUAL + write, opaque. A "step" is one frame: render-pass clear of
colour and depth + the draw. With --gpus N (torchrun, one rank per GPU) the frame is split sort-first into N
horizontal bands, vertex work replicated, and the finished bands are all-gathered over NCCL (scaling: strong).

One JSON line on stdout (rank 0). Keys beyond the base contract: roofline, cpu_baseline, e2e, clocks, gpu_launches,
plus gfragments_per_s / ms_per_frame breakdown.

  python bench.py                       # N=1, C3/M1
  python bench.py --impl reference      # the CPU arm: the oracle restatement of the reference path on host cores
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N
"""
import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from cpvulkan_b200 import build, capi, scenes  # noqa: E402

WIDTH, HEIGHT, NX, NY = 3840, 2160, 1000, 500


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f).get("hbm_gbs", 6650.0), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """SM clock and throttle reasons sampled DURING the timed region: through NVML every 2 ms (the timed region of the default run
    is a few milliseconds long), or — when the NVML binding is missing — by `nvidia-smi -lms 200`."""

    FIELDS = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
    NVML_REASONS = ((0x8, "hw_slowdown"), (0x40, "hw_thermal_slowdown"), (0x20, "sw_thermal_slowdown"), (0x4, "sw_power_cap"))

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.proc = index, [], None
        self.done = threading.Event()
        self.nvml_samples, self.nvml_max, self.nvml_mask = [], None, 0

    def _nvml_loop(self):
        import pynvml
        import torch
        pynvml.nvmlInit()
        p = torch.cuda.get_device_properties(self.index)
        try:
            h = pynvml.nvmlDeviceGetHandleByPciBusId(("%08x:%02x:%02x.0" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)).encode())
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
        self.nvml_max = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
        while not self.done.is_set():
            self.nvml_samples.append(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM))
            self.nvml_mask |= int(pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h))
            time.sleep(0.002)

    def run(self):
        try:
            self._nvml_loop()
            return
        except Exception:
            self.nvml_samples = []
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                self.rows.append([x.strip() for x in line.split(",")])
        except Exception:
            pass

    def mark(self):
        """Forget what was sampled so far: called right before the timed region starts."""
        self.nvml_samples.clear(); self.nvml_mask = 0; self.rows.clear()

    def stop(self):
        self.done.set()
        if self.proc:
            self.proc.terminate()
        if self.nvml_samples:
            self.join(timeout=1.0)
            reasons = sorted(name for bit, name in self.NVML_REASONS if self.nvml_mask & bit)
            return {"sm_mhz": float(statistics.median(self.nvml_samples)), "sm_max_mhz": float(self.nvml_max), "reasons": reasons,
                    "samples": len(self.nvml_samples), "source": "nvml, every 2 ms"}
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons, "samples": len(sm),
                "source": "nvidia-smi -lms 200"}


def oracle_draw_seconds(scene, repeats=1):
    """Time the oracle's equivalent of DrawIndexedCommand::Process (draw only; clears outside) on one host core."""
    lib = capi.load_oracle()
    mem = scenes.HostMemory()
    m = scenes.materialize(scene, mem.alloc)
    times, stats = [], capi.DrawStats()
    for _ in range(repeats):
        for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
            if img is not None and img.clear is not None:
                cv, is_ds = scenes.clear_value(img)
                lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds)
        t0 = time.perf_counter()
        rc = lib.cpvk_oracle_draw(C.byref(m.desc), C.byref(m.state), C.byref(stats))
        times.append(time.perf_counter() - t0)
        if rc != 0:
            raise RuntimeError(lib.cpvk_oracle_last_error().decode())
    return statistics.median(times), stats


def cpu_info():
    model = "unknown"
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    model = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    return model, os.cpu_count()


def run_reference(args):
    """--impl reference: the reference's own algorithm for the path on the host CPU. The reference ICD cannot be
    built in this image (LLVM-8 / Vulkan SDK / GSL / glm missing), so this arm times the oracle
    restatement, single-threaded like the reference's inline vkQueueSubmit (Queue.cpp:52-59)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build.build_oracle()
    scene = scenes.mesh_indexed(WIDTH, HEIGHT, NX, NY)
    for _ in range(min(args.warmup, 1)):
        oracle_draw_seconds(scene)
    t, st = oracle_draw_seconds(scene, repeats=max(1, min(args.steps, 5)))
    prims = 2 * NX * NY
    model, ncpu = cpu_info()
    val = prims / t / 1e6
    line = {
        "impl": "reference", "metric": "Mtris/s (vkCmdDrawIndexed, 1M triangles at 3840x2160, D32 depth test, opaque)", "value": val, "unit": "Mtris/s",
        "n_gpus": args.gpus, "steps": max(1, min(args.steps, 5)), "warmup": min(args.warmup, 1), "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "C3/M1: 1,000,000-triangle indexed grid, 3840x2160 RGBA8+D32, LESS_OR_EQUAL, opaque", "l2": "n/a (CPU)"},
        "gfragments_per_s": st.fragmentsCovered / t / 1e9,
        "cpu_baseline": {"value": val, "unit": "Mtris/s", "cores": 1, "kind": "port",
                         "sample": "full C3/M1 draw (1,000,000 triangles, %d fragments), draw only, median of runs; %s, %s logical CPUs; "
                                   "oracle omits the reference's JIT/indirect-call overhead (optimistic stand-in)" % (st.fragmentsCovered, model, ncpu)},
        "e2e": {"value": val, "unit": "Mtris/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), file=_RESULT_OUT, flush=True)


def bind_near_gpu(index):
    """Pin this process to the CPUs of the NUMA node GPU `index` hangs off (sysfs local_cpulist of its PCI function), so that
    the pinned host buffers of the e2e leg — allocated first-touch by this process — sit on the near side of the PCIe root
    complex; on a two-socket box the far side costs a third of the copy bandwidth. Returns what was done, for the record."""
    try:
        import torch
        p = torch.cuda.get_device_properties(index)
        path = "/sys/bus/pci/devices/%04x:%02x:%02x.0/local_cpulist" % (p.pci_domain_id, p.pci_bus_id, p.pci_device_id)
        with open(path) as f:
            text = f.read().strip()
        cpus = set()
        for part in text.split(","):
            if part:
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        near = cpus & allowed
        if not near or near == allowed:
            return "unchanged (one NUMA node or no topology information)"
        os.sched_setaffinity(0, near)
        return "NUMA-local: %d of %d CPUs" % (len(near), len(allowed))
    except Exception as e:  # no sysfs entry, no permission: measure as is
        return "unchanged (%s)" % type(e).__name__


def run_ours(args):
    import torch
    import torch.distributed as dist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torchrun for --gpus > 1")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the draw path has no CPU fallback")
    torch.cuda.set_device(local)
    affinity = bind_near_gpu(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from cpvulkan_b200.device import Device, SceneOnDevice

    build.build_cuda()
    scene = scenes.mesh_indexed(WIDTH, HEIGHT, NX, NY)
    # One explicit stream for everything: the draw kernels, the NCCL gather and the timing events. (torch's default
    # stream has handle 0, which the C ABI reads as "use the device's own stream" — events recorded on the default
    # stream would then not be ordered against the kernels.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    dev = Device(local, stream=stream.cuda_stream, stats=True)
    rows = HEIGHT // world
    band = (rank * rows, (rank + 1) * rows) if world > 1 else None

    # Colour attachment owned by torch. Multi-GPU: peer-mapped (symmetric memory) so that k_raster can store every
    # finished tile of this rank's band straight into the other GPUs' frames over NVLink — the gather of SURVEY §8(e)
    # fused into rasterisation; `--gather nccl` keeps the frame private and all-gathers the bands afterwards.
    symm = None
    if world > 1 and args.gather == "fused":
        import torch.distributed._symmetric_memory as symm_mem
        color_t = symm_mem.empty(scene.color.nbytes, dtype=torch.uint8, device=torch.device("cuda", local))
        color_t.zero_()
        symm = symm_mem.rendezvous(color_t, dist.group.WORLD)
    else:
        color_t = torch.zeros(scene.color.nbytes, dtype=torch.uint8, device="cuda")

    class Placed(SceneOnDevice):
        pass

    sod = SceneOnDevice.__new__(Placed)
    sod.dev, sod.scene, sod.owned = dev, scene, []

    def alloc(name, nbytes, init):
        if name == "color":
            return color_t.data_ptr()
        a = dev.alloc(nbytes)
        sod.owned.append(a)
        if init is not None:
            dev.upload(a, np.ascontiguousarray(init).view(np.uint8).reshape(-1)[:nbytes])
        return a

    sod.m = scenes.materialize(scene, alloc)
    sod.pipeline = dev.create_pipeline(sod.m.desc)
    sod.m.state.pipeline = sod.pipeline.value
    if band:
        sod.m.state.bandY0, sod.m.state.bandY1 = band
    band_bytes = rows * scene.color.pitch
    if symm is not None:
        peers = [int(p) for i, p in enumerate(symm.buffer_ptrs) if i != rank]
        sod.m.state.mirrorCount = len(peers)
        for i, ptr in enumerate(peers):
            sod.m.state.mirrorColor0[i] = ptr

    def frame():
        sod.clear(band_only=world > 1)
        sod.draw()
        if symm is not None:
            symm.barrier()  # every GPU's tiles have landed in every frame before anyone reads or redraws
        elif world > 1:
            dist.all_gather_into_tensor(color_t, color_t[rank * band_bytes:(rank + 1) * band_bytes])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(max(args.warmup, 3)):
        frame()
    barrier()
    st = dev.stats()
    n_cov, n_pass = int(st.fragmentsCovered), int(st.fragmentsWritten)
    if world > 1:
        # every rank must now hold the same, complete frame (cheap integrity check of the exchange, outside the timed region)
        digest = torch.stack([color_t.view(torch.int32).to(torch.int64).sum(), (color_t.view(torch.int32)[::4097].to(torch.int64) * 31).sum()])
        all_digests = [torch.zeros_like(digest) for _ in range(world)]
        dist.all_gather(all_digests, digest)
        if any(not torch.equal(all_digests[0], x) for x in all_digests):
            raise SystemExit("ranks disagree on the gathered frame")
        # ... and that frame must be, byte for byte, what one GPU renders without bands
        whole = torch.zeros_like(color_t)
        saved = (sod.m.state.bandY0, sod.m.state.bandY1, sod.m.state.mirrorCount, sod.m.state.color[0].address, sod.m.color_attachment.address)
        sod.m.state.bandY0 = sod.m.state.bandY1 = 0
        sod.m.state.mirrorCount = 0
        sod.m.state.color[0].address = sod.m.color_attachment.address = whole.data_ptr()
        sod.clear(); sod.draw(); dev.sync()
        sod.m.state.bandY0, sod.m.state.bandY1, sod.m.state.mirrorCount, sod.m.state.color[0].address, sod.m.color_attachment.address = saved
        if not torch.equal(whole, color_t):
            raise SystemExit("the gathered frame differs from the single-GPU frame")
        del whole
    if world > 1:
        t = torch.tensor([n_cov, n_pass], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_cov, n_pass = int(t[0]), int(t[1])
    dev.set_stats(False)

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    launches0 = dev.launch_count()
    barrier()
    if sampler and sampler.nvml_samples:
        sampler.mark()  # NVML samples every 2 ms: keep only those taken inside the timed region (nvidia-smi's 200 ms rows are all kept)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        frame()
    e1.record()
    barrier()
    clocks = sampler.stop() if sampler else None
    ms_total = e0.elapsed_time(e1)
    launches = dev.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t[0])
    ms_step = ms_total / args.steps
    prims = 2 * NX * NY

    # per-kernel durations (CUDA events on the launching stream, second pass so the events do not perturb `value`)
    dev.set_timing(True)
    vs, su, bn, rs = [], [], [], []
    for _ in range(args.steps):
        sod.clear(band_only=world > 1)
        sod.draw()
        s2 = dev.stats()
        vs.append(s2.msVertex); su.append(s2.msSetup); bn.append(s2.msBin); rs.append(s2.msRaster)
    dev.set_timing(False)
    ms_raster = statistics.mean(rs)
    bin_entries = int(s2.binEntries)

    # roofline of the dominant kernel (cpvk_k_raster): algorithmic attachment bytes per launch, SURVEY §8(d):
    #   N_cov * bD (depth read) + N_pass * (bD + bC) (depth write + colour write); RGBA8 + D32 -> 12 B / fragment
    local_cov, local_pass = int(st.fragmentsCovered), int(st.fragmentsWritten)
    b_alg = local_cov * 4 + local_pass * (4 + 4)
    peak, peak_src = measured_peaks()
    achieved = b_alg / (ms_raster * 1e-3) / 1e9 if ms_raster > 0 else 0.0

    # e2e on N > 1 GPUs (every rank takes part): each frame, every rank uploads the (replicated) geometry from its own pinned
    # buffers over its own PCIe link, renders its band — the fused gather and its barrier stay in the frame — and reads ITS band
    # back into pinned host memory: the host ends up with the whole frame, one band per process, 1/N of the read-back per link.
    # The read-back of frame k overlaps the upload and rendering of frame k+1: the finished band is copied to one of two
    # staging buffers on the draw stream (16 MB, device to device) and a second device object, on its own stream, carries
    # it to the host; events order the two streams in both directions.
    e2e_multi = None
    if world > 1:
        names = ["vb", "ib", "ubo"]
        staged = {}
        for nme in names:
            data = scene.buffers[nme]
            a = dev.alloc(data.nbytes, host_shadow=True)
            dev.shadow(a)[:data.nbytes] = data
            staged[nme] = (a, data.nbytes)
        stream2 = torch.cuda.Stream()
        dev2 = Device(local, stream=stream2.cuda_stream, stats=False)
        band_addr = color_t.data_ptr() + rank * band_bytes
        slots = []
        for _ in range(2):
            staging = torch.empty(band_bytes, dtype=torch.uint8, device="cuda")
            out_dev = dev2.alloc(band_bytes, host_shadow=True)  # only its pinned shadow is used, as the read-back target
            slots.append({"staging": staging, "host": dev2.allocs[out_dev][1], "copied": torch.cuda.Event(), "read": torch.cuda.Event()})
            slots[-1]["read"].record(stream2)

        def e2e_frame(k):
            sl = slots[k % 2]
            for nme in names:
                src_alloc, nbytes = staged[nme]
                dev.upload_async(sod.m.addr[nme], dev.allocs[src_alloc][1], nbytes)
            frame()
            stream.wait_event(sl["read"])                       # the staging buffer's previous contents have reached the host
            dev.copy_rows(sl["staging"].data_ptr(), band_bytes, band_addr, band_bytes, band_bytes, 1)
            sl["copied"].record(stream)
            stream2.wait_event(sl["copied"])
            dev2.download_into_async(sl["host"], sl["staging"].data_ptr(), band_bytes)
            sl["read"].record(stream2)

        for k in range(4):
            e2e_frame(k)
        torch.cuda.synchronize()
        want = color_t[rank * band_bytes:(rank + 1) * band_bytes].cpu()
        for sl in slots:
            got = torch.from_numpy(np.ctypeslib.as_array(C.cast(sl["host"], C.POINTER(C.c_uint8)), shape=(band_bytes,)).copy())
            if not torch.equal(got, want):
                raise SystemExit("e2e read-back of rank %d's band differs from the resident frame" % rank)
        barrier()
        t0 = time.perf_counter()
        for k in range(args.steps):
            e2e_frame(k)
        barrier()  # torch.cuda.synchronize() inside: both streams have drained
        t = torch.tensor([(time.perf_counter() - t0) * 1e3 / args.steps], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms = float(t[0])
        e2e_multi = {"value": prims / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": world * sum(v[1] for v in staged.values()),
                     "d2h_bytes_per_step": world * band_bytes, "ms_per_step": e2e_ms,
                     "through": "C ABI on every rank (cpvk_cuda_mem_upload of the replicated geometry / clear / draw with the fused gather / copy_rows of "
                                "the rank's band to a staging buffer / mem_download_async on a second device object), pinned host buffers, "
                                "read-back of frame k overlapped with frame k+1; max over ranks"}
        dev2.close()

    line = None
    if rank == 0:
        # e2e through the C ABI with HOST buffers: every frame uploads its inputs (vertices, indices, uniforms) from pinned
        # host memory, clears, draws and reads the colour result back into pinned host memory. Like a double-buffered
        # application, two device objects (each with its own stream, scratch and frame) alternate frames, so the read-back of
        # frame k overlaps the upload and rendering of frame k+1; a frame's buffers are reused only after its read-back is done.
        e2e = e2e_multi
        if world == 1:
            names = ["vb", "ib", "ubo"]
            lanes = []
            n_lanes = int(os.environ.get("CPVK_E2E_LANES", "2"))
            for i in range(n_lanes):
                ldev = dev if i == 0 else Device(local, stats=False)  # the second one runs on its own stream
                lsod = sod if i == 0 else SceneOnDevice(ldev, scene)
                staged = {}
                for nme in names:
                    data = scene.buffers[nme]
                    a = ldev.alloc(data.nbytes, host_shadow=True)
                    ldev.shadow(a)[:data.nbytes] = data
                    staged[nme] = (a, data.nbytes)
                out_dev = ldev.alloc(scene.color.nbytes, host_shadow=True)  # only its pinned shadow is used as the readback target
                lanes.append((ldev, lsod, staged, ldev.allocs[out_dev][1]))
            h2d = sum(v[1] for v in lanes[0][2].values())
            d2h = scene.color.nbytes

            def e2e_step(k):
                ldev, lsod, staged, out_host = lanes[k % n_lanes]
                ldev.sync()  # frame k-2 (same lane) has been read back: its buffers are free
                for nme in names:
                    src_alloc, nbytes = staged[nme]
                    ldev.upload_async(lsod.m.addr[nme], ldev.allocs[src_alloc][1], nbytes)
                lsod.clear()
                lsod.draw()
                ldev.download_into_async(out_host, lsod.m.addr["color"], d2h)

            for k in range(2 * n_lanes):
                e2e_step(k)
            for lane in lanes:
                lane[0].sync()
            # the read-back frame must be the frame the resident path produced
            ref_frame = color_t.cpu().numpy()
            for lane in lanes:
                got = np.ctypeslib.as_array(C.cast(lane[3], C.POINTER(C.c_uint8)), shape=(d2h,))
                if not np.array_equal(got, ref_frame):
                    raise SystemExit("e2e read-back differs from the resident frame")
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            for k in range(args.steps):
                e2e_step(k)
            for lane in lanes:
                lane[0].sync()
            e2e_ms = (time.perf_counter() - t0) * 1e3 / args.steps
            e2e = {"value": prims / (e2e_ms * 1e-3) / 1e6, "unit": "Mtris/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms,
                   "through": "C ABI (cpvk_cuda_mem_upload / clear / draw / mem_download_async + sync) with pinned host buffers; two device objects alternate frames (double buffering)"}
            for lane in lanes[1:]:
                lane[1].close()
                lane[0].close()
        extras = None
        if world == 1 and not args.no_extras:
            extras = secondary_configs(dev, torch)
        cpu = None
        if world == 1 and not args.no_cpu:
            build.build_oracle()
            t_cpu, st_cpu = oracle_draw_seconds(scene, repeats=3)
            model, ncpu = cpu_info()
            cpu = {"value": prims / t_cpu / 1e6, "unit": "Mtris/s", "cores": 1, "kind": "port",
                   "sample": "full C3/M1 draw, draw only, median of 3; %s, %d logical CPUs; optimistic stand-in for the reference ICD (no JIT/indirection overhead)" % (model, ncpu),
                   "fragments_match_gpu": int(st_cpu.fragmentsCovered) == n_cov}
        line = {
            "metric": "Mtris/s (vkCmdDrawIndexed, 1M triangles at 3840x2160, D32 depth test, opaque)", "value": prims / (ms_step * 1e-3) / 1e6, "unit": "Mtris/s",
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "C3/M1: 1,000,000-triangle indexed grid, 3840x2160 RGBA8+D32, LESS_OR_EQUAL, opaque; step = clear + draw" +
                                   ((" + bands stored into the peers' frames by k_raster over NVLink (fused gather) + barrier" if symm is not None else " + NCCL all-gather of %d bands" % world) if world > 1 else ""),
                       "parallelism": "sort-first bands x%d" % world,
                       "l2": "working set (indices 12 MB + vertices 16 MB + shaded vertices 16 MB + setup records 104 MB + tile lists 5 MB + targets 66 MB) exceeds the 126 MB L2; no explicit flush",
                       "cpu_affinity": affinity},
            "gfragments_per_s": n_cov / (ms_step * 1e-3) / 1e9, "ms_per_frame": ms_step,
            "fragments_covered": n_cov, "fragments_written": n_pass, "bin_entries_rank0": bin_entries,
            "kernel_ms_rank0": {"vertex": statistics.mean(vs), "setup": statistics.mean(su), "bin": statistics.mean(bn), "raster": ms_raster},
            "roofline": {"bound": "hbm", "kernel": "cpvk_k_raster", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg, "traffic": raster_traffic(),
                         "note": "front end of C3/M1 is instruction-bound (8 px/triangle); see DESIGN.md"},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": launches,
        }
        if extras:
            line["other_configs"] = extras
    sod.close()
    dev.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=_RESULT_OUT, flush=True)


def raster_traffic():
    """DRAM bytes of one cpvk_k_raster launch from the committed `ncu --set full` capture (never measured inside a bench run)."""
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "raster_traffic.json")) as f:
            t = json.load(f)
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        return None


def secondary_configs(dev, torch):
    """Short measurements of the other BASELINE configs on the same device (reduced so the default run stays short):
    C4 = blended, LINEAR-filtered full-screen quads at 7680x4320 RGBA16F (40 of the 2,000 quads);
    C5 = vkCmdBlitImage / vkCmdCopyImage at 7680x4320. CUDA events on the launching stream, inputs resident."""
    from cpvulkan_b200 import capi
    from cpvulkan_b200.device import SceneOnDevice
    peak, _ = measured_peaks()
    out = {}

    def timed(fn, reps):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    quads = 40
    sc = scenes.overdraw_quads(width=7680, height=4320, quads=quads, tex_size=1024)
    s = SceneOnDevice(dev, sc)
    dev.set_stats(True)
    s.render()
    frags = int(dev.stats().fragmentsCovered)
    dev.set_stats(False)
    ms = timed(s.render, 3)
    out["C4_overdraw"] = {"workload": "%d of the 2,000 alpha-blended LINEAR-textured full-screen quads, 7680x4320 RGBA16F; step = clear + draw" % quads,
                          "ms_per_step": ms, "gfragments_per_s": frags / (ms * 1e-3) / 1e9, "mtris_per_s": 2 * quads / (ms * 1e-3) / 1e6,
                          "roofline_frac": frags * 16 / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_fragment": 16}
    s.close()

    W, H = 7680, 4320

    def image(fmt, w, h, texel):
        t = torch.randint(0, 255, (w * h * texel,), dtype=torch.uint8, device="cuda")
        return t, capi.Attachment(t.data_ptr(), w, h, w * texel, fmt)

    s8, a8 = image(37, W, H, 4)
    d16, a16 = image(97, W, H, 8)
    s4, a4 = image(37, W // 2, H // 2, 4)
    d8, _ = image(37, W, H, 4)
    b1 = capi.Blit(a8, a16, 0, 0, W, H, 0, 0, W, H, 0)
    b2 = capi.Blit(a4, a16, 0, 0, W // 2, H // 2, 0, 0, W, H, 1)
    for key, fn, nbytes in (("blit_8k_rgba8_to_rgba16f_nearest", lambda: dev.blit(b1), W * H * 12),
                            ("blit_4k_to_8k_rgba16f_linear", lambda: dev.blit(b2), W * H * 8 + W * H),
                            ("copy_image_8k_rgba8", lambda: dev.copy_rows(d8.data_ptr(), W * 4, s8.data_ptr(), W * 4, W * 4, H), W * H * 8)):
        ms = timed(fn, 5)
        out["C5_" + key] = {"ms": ms, "GBps": nbytes / ms / 1e6, "roofline_frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes": nbytes}
    return out


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="fused", choices=["fused", "nccl"], help="multi-GPU exchange: peer stores from k_raster (default) or an NCCL all-gather after the draw")
    ap.add_argument("--no-extras", action="store_true", help="skip the short C4 / C5 measurements reported under other_configs")
    args = ap.parse_args()
    # stdout carries exactly ONE line, the JSON record: whatever libraries print to file descriptor 1 on the way (NCCL's
    # version banner under torchrun, for one) is sent to stderr instead
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()

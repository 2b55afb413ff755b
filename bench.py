#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 draw path (BASELINE.json: Mtris/s and Gfragments/s for
vkCmdDrawIndexed at 4K; ms/frame at 1/2/4/8 B200).

Default workload (config.workload "C3/M1"): one vkCmdDrawIndexed of a 1000x500-quad grid (1,000,000
triangles, 3,000,000 u32 indices, 501,501 vertices {vec4 pos, vec4 rgba}) at 3840x2160, R8G8B8A8_UNORM + D32_SFLOAT,
depth LESS_OR_EQUAL + write, opaque. `--config c4` runs BASELINE config 4 instead: all 2,000 alpha-blended
LINEAR-textured full-screen quads at 7680x4320 RGBA16F. A "step" is one frame: render-pass clear + the draw
(+ the band exchange on N > 1 GPUs).

With --gpus N (torchrun, one rank per GPU) the frame is split sort-first into N bands of tile rows, vertex work is
replicated, and k_raster stores every finished tile of a rank's band straight into the presenting GPU's frame (rank 0;
`--gather all`: into every GPU's frame) over NVLink — peer memory mapped through the C ABI's cudaIpc export / import, the
handles exchanged with torch.distributed; cpvk_cuda_peer_barrier on the draw stream (flag words in the same peer-mapped
memory; `--barrier nccl`: a 4-byte NCCL all-reduce) orders the GPUs after the stores.
Scaling is strong (fixed frame).

Timing: the K steps asked for are one block, bracketed by barrier + synchronize and timed with CUDA events on the launching
stream; blocks are repeated until at least 0.5 s and 200 frames have been timed (>= 5 blocks; a workload whose frames take
seconds stops after 5 blocks), and `value` comes from the
MEDIAN block (max over ranks per block), so a scheduling hiccup in one block does not move the headline.

One JSON line on stdout (rank 0). Keys beyond the base contract: roofline, cpu_baseline, e2e, clocks, gpu_launches, icd,
blocks, kernel_ms_rank0, other_configs.

  python bench.py                       # N=1, C3/M1
  python bench.py --config c4           # N=1, the full C4 frame
  python bench.py --impl reference      # the CPU arm: the oracle restatement of the reference path on the host cores
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


TILE = 32


class Workload:
    """What one step renders, and how its numbers are named."""

    def __init__(self, key, quads=None):
        self.key = key
        if key == "c3":
            self.scene = scenes.mesh_indexed(WIDTH, HEIGHT, NX, NY)
            self.prims = 2 * NX * NY
            self.metric = "Mtris/s (vkCmdDrawIndexed, 1M triangles at 3840x2160, D32 depth test, opaque)"
            self.unit, self.scale = "Mtris/s", 1e6
            self.workload = "C3/M1: 1,000,000-triangle indexed grid, 3840x2160 RGBA8+D32, LESS_OR_EQUAL, opaque; step = clear + draw"
            self.l2 = ("working set (indices 12 MB + vertices 16 MB + shaded vertices 16 MB + setup records 104 MB + tile lists 5 MB + "
                       "targets 66 MB) exceeds the 126 MB L2; no explicit flush")
        else:
            self.quads = quads or int(os.environ.get("CPVK_BENCH_C4_QUADS", "2000"))  # the environment knob is for tuning runs only
            self.scene = scenes.overdraw_quads(7680, 4320, quads=self.quads, tex_size=1024)
            self.prims = 2 * self.quads
            self.metric = "Gfragments/s (%s alpha-blended LINEAR-textured full-screen quads at 7680x4320 RGBA16F)" % ("2,000" if self.quads == 2000 else str(self.quads))
            self.unit, self.scale = "Gfragments/s", 1e9
            self.workload = "C4: %d alpha-blended LINEAR / REPEAT textured full-screen quads, 7680x4320 RGBA16F, 1024^2 RGBA8 texture; step = clear + draw" % self.quads
            self.l2 = "the 265 MB colour target exceeds the 126 MB L2; tiles live in shared memory between the quads; no explicit flush"

    def units(self, n_cov):
        """What `value` counts per step: triangles (C3) or coverage-passing fragments (C4)."""
        return self.prims if self.key == "c3" else n_cov

    def algorithmic_bytes(self, n_cov, n_pass):
        # depth read per covered fragment, depth write + colour write (+ colour read when blending) per passing one
        return n_cov * 4 + n_pass * 8 if self.key == "c3" else n_pass * 16


def band_rows(height, rank, world):
    """Rows [y0, y1) of rank's sort-first band: whole 32-row tile rows, spread evenly (the same split cpvk_cuda_draw makes for
    the members of a group), so no tile is rasterised by two GPUs."""
    tile_rows = (height + TILE - 1) // TILE
    a, b = rank * tile_rows // world, (rank + 1) * tile_rows // world
    return min(a * TILE, height), min(b * TILE, height)


def oracle_draw_seconds(scene, threads=1, repeats=1):
    """Time the oracle's equivalent of DrawIndexedCommand::Process (draw only; clears outside). threads > 1: the frame is cut
    into that many horizontal windows rendered concurrently (cpvk_oracle_draw_window; pixels are independent in the reference,
    every window replays the vertex stage) — the reference itself is single-threaded (Queue.cpp:52-59)."""
    from concurrent.futures import ThreadPoolExecutor
    lib = capi.load_oracle()
    mem = scenes.HostMemory()
    m = scenes.materialize(scene, mem.alloc)
    times, covered = [], 0
    h, w = scene.color.height, scene.color.width
    cuts = [(i * h // threads, (i + 1) * h // threads) for i in range(threads)]

    def one(window):
        st = capi.DrawStats()
        rc = lib.cpvk_oracle_draw_window(C.byref(m.desc), C.byref(m.state), 0, window[0], w, window[1], C.byref(st))
        if rc != 0:
            raise RuntimeError(lib.cpvk_oracle_last_error().decode())
        return int(st.fragmentsCovered)

    with ThreadPoolExecutor(max_workers=threads) as pool:
        for _ in range(repeats):
            for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
                if img is not None and img.clear is not None:
                    cv, is_ds = scenes.clear_value(img)
                    lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds)
            t0 = time.perf_counter()
            covered = sum(pool.map(one, cuts))
            times.append(time.perf_counter() - t0)
    return statistics.median(times), covered


def cpu_sample(work, threads):
    """The bounded CPU sample of a workload: C3 = the whole draw; C4 = the whole 8K frame for 2 of the quads (the cost is linear
    in quads: every quad covers every pixel)."""
    if work.key == "c3":
        return work.scene, 1.0, "full C3/M1 draw (1,000,000 triangles), draw only"
    q = 2
    return scenes.overdraw_quads(7680, 4320, quads=q, tex_size=1024), work.quads / q, "%d of the %d quads at 7680x4320 (cost is linear in quads), draw only" % (q, work.quads)


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
    """--impl reference: the reference's own algorithm for the path on the host CPU. The reference ICD cannot be built in this
    image (LLVM-8 / Vulkan SDK / GSL / glm >= 0.9.9 missing), so this arm times the oracle restatement (pinned
    against the reference's own Draw.cpp / ImageSampler.cpp / Formats.cpp) on every host core it can use."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    build.build_oracle()
    work = Workload(args.config)
    threads = max(1, len(os.sched_getaffinity(0)))
    scene, factor, what = cpu_sample(work, threads)
    for _ in range(args.warmup):
        oracle_draw_seconds(scene, threads)
    times = []
    covered = 0
    for _ in range(args.steps):
        t, covered = oracle_draw_seconds(scene, threads)
        times.append(t * factor)
    t = statistics.median(times)
    t1, _ = oracle_draw_seconds(scene, 1)
    model, ncpu = cpu_info()
    units = work.units(int(covered * factor))
    val = units / t / work.scale
    line = {
        "impl": "reference", "metric": work.metric, "value": val, "unit": work.unit,
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3, "higher_is_better": True,
        "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": work.workload, "parallelism": "sort-first bands x%d" % args.gpus, "l2": "n/a (CPU)"},
        "gfragments_per_s": covered * factor / t / 1e9, "ms_per_frame": t * 1e3,
        "cpu_baseline": {"value": val, "unit": work.unit, "cores": threads, "kind": "port",
                         "sample": "%s, median of %d steps, %d threads (one horizontal window each); one thread: %.3f %s; %s, %s logical CPUs; "
                                   "the oracle omits the reference's JIT / indirect-call overhead (optimistic stand-in)"
                                   % (what, args.steps, threads, units / (t1 * factor) / work.scale, work.unit, model, ncpu)},
        "e2e": {"value": val, "unit": work.unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
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


class Rig:
    """One rank's device, scene and frame function for a workload; on N > 1 ranks also the peer mapping of the colour frame."""

    def __init__(self, work, torch, dist, stream, local, rank, world, gather, barrier="flags"):
        from cpvulkan_b200.device import Device, SceneOnDevice
        self.torch, self.dist, self.work, self.rank, self.world = torch, dist, work, rank, world
        self.dev = Device(local, stream=stream.cuda_stream, stats=True)
        # front-end overlap across frames (include/cpvk_cuda.h): the bench's own stream carries nothing between two frames that
        # writes a frame's inputs except where it calls flush() (the e2e leg on N > 1 GPUs); CPVK_OVERLAP=0 switches it off
        if os.environ.get("CPVK_OVERLAP", "1") != "0":
            self.dev.set_overlap(True)
        self.sod = SceneOnDevice(self.dev, work.scene)
        scene = work.scene
        self.band = band_rows(scene.color.height, rank, world) if world > 1 else None
        self.peers = []
        self.token = torch.zeros(1, dtype=torch.int32, device="cuda") if world > 1 else None
        if world > 1:
            st = self.sod.m.state
            st.bandY0, st.bandY1 = self.band
            # the colour frame of every rank, mapped into this process (cudaIpc through the C ABI; torch.distributed only carries
            # the 64-byte handles)
            handles = [None] * world
            dist.all_gather_object(handles, self.dev.export_handle(self.sod.m.addr["color"]))
            targets = [r for r in range(world) if r != rank] if gather == "all" else ([0] if rank != 0 else [])
            for i, r in enumerate(targets):
                addr = self.dev.import_handle(handles[r])
                self.peers.append(addr)
                st.mirrorColor0[i] = addr
            st.mirrorCount = len(targets)
            # the per-frame ordering step between the GPUs: cpvk_cuda_peer_barrier on flag words every rank maps (default), or a
            # 4-byte NCCL all-reduce (--barrier nccl)
            self.barrier, self.sequence, self.flag_arrays = barrier, 0, []
            if barrier == "flags":
                mine = self.dev.alloc(64)
                self.dev.sync()  # the zero fill of the allocation is done before any peer can signal into it
                dist.all_gather_object(handles, self.dev.export_handle(mine))
                for r in range(world):
                    if r == rank:
                        self.flag_arrays.append(mine)
                    else:
                        addr = self.dev.import_handle(handles[r])
                        self.peers.append(addr)
                        self.flag_arrays.append(addr)
                dist.barrier()

    def frame(self):
        self.sod.clear(band_only=self.world > 1)
        self.sod.draw()
        if self.world > 1:
            # orders the GPUs: the collective starts on a rank after its raster kernel (and its peer stores) finished, and ends
            # everywhere only after it started everywhere
            if self.barrier == "flags":
                self.sequence += 1
                self.dev.peer_barrier(self.flag_arrays, self.rank, self.sequence)
            else:
                self.dev.flush()  # foreign work behind a draw on this stream: the draw is validated (and replayed if need be) first
                self.dist.all_reduce(self.token)

    def single_gpu_frame(self):
        """The frame one GPU renders without bands, into a second image (the reference for the exchange)."""
        st = self.sod.m.state
        saved = (st.bandY0, st.bandY1, st.mirrorCount, st.color[0].address, self.sod.m.color_attachment.address)
        whole = self.dev.alloc(self.work.scene.color.nbytes)
        st.bandY0 = st.bandY1 = 0
        st.mirrorCount = 0
        st.color[0].address = self.sod.m.color_attachment.address = whole
        self.sod.clear(); self.sod.draw(); self.dev.sync()
        out = self.dev.download(whole, self.work.scene.color.nbytes)
        st.bandY0, st.bandY1, st.mirrorCount, st.color[0].address, self.sod.m.color_attachment.address = saved
        self.dev.free(whole)
        return out

    def close(self):
        self.dev.sync()
        for a in self.peers:
            self.dev.unimport(a)
        self.sod.close()
        self.dev.close()


HOST_ISSUE_MS = []


def timed_blocks(torch, dist, world, frame, steps, flush, min_seconds=0.5, min_frames=200, min_blocks=5, max_blocks=400, enough_seconds=3.0):
    """Blocks of exactly `steps` frames, each bracketed by barrier + synchronize and timed with CUDA events on the current
    stream; per block the max over ranks. Returns the list of block times in ms."""
    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    blocks, total_ms, frames = [], 0.0, 0
    # at least min_blocks blocks and min_seconds; then on until min_frames frames, unless enough_seconds have been timed already
    while len(blocks) < max_blocks and (len(blocks) < min_blocks or total_ms < min_seconds * 1e3 or (frames < min_frames and total_ms < enough_seconds * 1e3)):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        t_host = time.perf_counter()
        for _ in range(steps):
            frame()
        HOST_ISSUE_MS.append((time.perf_counter() - t_host) * 1e3 / steps)  # how long the host needs to issue a frame (it may wait for the GPU inside)
        flush()  # the library may still owe the last draw its validation (include/cpvk_cuda.h, set_overlap): nothing foreign — this event — goes behind it before that
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t[0])
        blocks.append(ms); total_ms += ms; frames += steps
    return blocks


def measure(rig, steps, warmup, sampler=None):
    """Resident throughput of rig.frame(): -> dict(ms_step, blocks, n_cov, n_pass, local_cov, local_pass, launches_per_step, kernel_ms)."""
    torch, dist, world, dev = rig.torch, rig.dist, rig.world, rig.dev
    for _ in range(max(warmup, 3)):
        rig.frame()
    dev.flush()  # synchronize() below is not a call of the library: the last draw is validated (and replayed, with its barrier) first
    torch.cuda.synchronize()
    st = dev.stats()
    local_cov, local_pass, bin_entries = int(st.fragmentsCovered), int(st.fragmentsWritten), int(st.binEntries)
    n_cov, n_pass = local_cov, local_pass
    if world > 1:
        t = torch.tensor([n_cov, n_pass], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_cov, n_pass = int(t[0]), int(t[1])
    dev.set_stats(False)
    launches0 = dev.launch_count()
    if sampler and sampler.nvml_samples:
        sampler.mark()  # NVML samples every 2 ms: keep only those taken inside the timed region
    del HOST_ISSUE_MS[:]
    blocks = timed_blocks(torch, dist, world, rig.frame, steps, rig.dev.flush)
    host_issue_ms = statistics.median(HOST_ISSUE_MS)
    launches = (dev.launch_count() - launches0) / (len(blocks) * steps)
    ms_step = statistics.median(blocks) / steps
    # per-kernel durations (CUDA events inside the library, a separate pass so that they do not perturb `value`)
    dev.set_timing(True)
    vs, su, bn, rs = [], [], [], []
    for _ in range(min(steps, 10)):
        rig.sod.clear(band_only=world > 1)
        rig.sod.draw()
        s2 = dev.stats()
        vs.append(s2.msVertex); su.append(s2.msSetup); bn.append(s2.msBin); rs.append(s2.msRaster)
    dev.set_timing(False)
    if world > 1:
        dist.all_reduce(rig.token)
        torch.cuda.synchronize()
    return {"ms_step": ms_step, "blocks": blocks, "n_cov": n_cov, "n_pass": n_pass, "local_cov": local_cov, "local_pass": local_pass,
            "launches_per_step": launches, "bin_entries": bin_entries, "host_issue_ms": host_issue_ms,
            "kernel_ms": {"vertex": statistics.mean(vs), "setup": statistics.mean(su), "bin": statistics.mean(bn), "raster": statistics.mean(rs)}}


def verify_exchange(rig):
    """N > 1: the frame rank 0 holds after the exchange must be, byte for byte, the frame one GPU renders without bands."""
    torch = rig.torch
    rig.frame()
    rig.dev.flush()  # before work that is not the library's: a first frame usually does not fit the launch plan and is replayed here
    torch.cuda.synchronize()
    rig.dist.barrier()
    if rig.rank == 0:
        got = rig.dev.download(rig.sod.m.addr["color"], rig.work.scene.color.nbytes)
        want = rig.single_gpu_frame()
        if not np.array_equal(got, want):
            raise SystemExit("the gathered frame differs from the single-GPU frame (%d bytes)" % int((got != want).sum()))
    rig.dist.barrier()


def link_floor_ms(torch, h2d, d2h, reps=20):
    """What this box's host link allows for an e2e frame: plain pinned-memory copies of the frame's upload and read-back sizes,
    both directions at once on two streams, nothing else running (ms per pair). e2e's ms_per_step cannot go below it."""
    hu = torch.empty(h2d, dtype=torch.uint8).pin_memory(); du = torch.empty(h2d, dtype=torch.uint8, device="cuda")
    hd = torch.empty(d2h, dtype=torch.uint8).pin_memory(); dd = torch.empty(d2h, dtype=torch.uint8, device="cuda")
    s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

    def pair():
        with torch.cuda.stream(s1):
            du.copy_(hu, non_blocking=True)
        with torch.cuda.stream(s2):
            hd.copy_(dd, non_blocking=True)

    for _ in range(3):
        pair()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        pair()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) * 1e3 / reps


def e2e_single(torch, rig, local, steps):
    """e2e through the C ABI with HOST buffers on one GPU: every frame uploads its inputs (vertices, indices, uniforms) from pinned
    host memory, clears, draws and reads the colour result back into pinned host memory. Like a double-buffered application,
    two device objects (each with its own stream, scratch and frame) alternate frames, so the read-back of frame k overlaps the
    upload and rendering of frame k+1; a frame's buffers are reused only after its read-back is done."""
    from cpvulkan_b200.device import Device, SceneOnDevice
    scene = rig.work.scene
    names = [n for n in ("vb", "ib", "ubo") if n in scene.buffers]
    lanes = []
    n_lanes = max(2, int(os.environ.get("CPVK_E2E_LANES", "3")))  # frame k is submitted before frame k - 1 is collected: two frames at least
    for i in range(n_lanes):
        ldev = Device(local, stats=False)  # runs on its own stream
        lsod = SceneOnDevice(ldev, scene)
        staged = {}
        for nme in names:
            data = scene.buffers[nme]
            a = ldev.alloc(data.nbytes, host_shadow=True)
            ldev.shadow(a)[:data.nbytes] = data
            staged[nme] = (a, data.nbytes)
        out_dev = ldev.alloc(scene.color.nbytes, host_shadow=True)  # only its pinned shadow is used as the read-back target
        lanes.append((ldev, lsod, staged, ldev.allocs[out_dev][1]))
    h2d = sum(v[1] for v in lanes[0][2].values())
    d2h = scene.color.nbytes

    def submit(k):
        ldev, lsod, staged, out_host = lanes[k % n_lanes]
        ldev.sync()  # frame k - n_lanes (same lane) has been read back: its buffers are free
        for nme in names:
            src_alloc, nbytes = staged[nme]
            ldev.upload_async(lsod.m.addr[nme], ldev.allocs[src_alloc][1], nbytes)
        lsod.clear()
        lsod.draw()

    def collect(k):
        ldev, lsod, staged, out_host = lanes[k % n_lanes]
        ldev.download_into_async(out_host, lsod.m.addr["color"], d2h)

    def run(count):
        # software pipelining as an application with frames in flight does it: submit frame k (upload + clear + draw), then
        # ask for frame k - 1's pixels. The read-back call has to look at frame k - 1's binning verdict first (speculative
        # draws, include/cpvk_cuda.h) — by now that frame's upload is done and the call does not stall the host behind it,
        # so frame k's upload and frame k - 1's read-back use the two directions of the link at the same time.
        for k in range(count):
            submit(k)
            if k > 0:
                collect(k - 1)
        collect(count - 1)
        for lane in lanes:
            lane[0].sync()

    run(2 * n_lanes)
    ref_frame = rig.dev.download(rig.sod.m.addr["color"], d2h)  # the read-back frame must be the frame the resident path produced
    for lane in lanes:
        got = np.ctypeslib.as_array(C.cast(lane[3], C.POINTER(C.c_uint8)), shape=(d2h,))
        if not np.array_equal(got, ref_frame):
            raise SystemExit("e2e read-back differs from the resident frame")
    n = max(steps, min(200, int(0.5 / 1e-3)))
    if rig.work.key != "c3":
        n = steps
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n)
    ms = (time.perf_counter() - t0) * 1e3 / n
    for lane in lanes:
        lane[1].close()
        lane[0].close()
    return {"ms_per_step": ms, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "frames_timed": n,
            "link_floor_ms": link_floor_ms(torch, h2d, d2h),
            "through": "C ABI (cpvk_cuda_mem_upload / clear / draw / mem_download_async + sync) with pinned host buffers; %d device objects alternate frames, frame k is submitted before frame k-1 is read back (frames in flight)" % n_lanes}


def e2e_multi(torch, dist, rig, local, steps):
    """e2e on N > 1 GPUs (every rank takes part): per frame each rank uploads 1/N of the geometry from its own pinned buffer over
    its own PCIe link, an NCCL all-gather over NVLink completes every rank's copy (vertex work is replicated, so every GPU needs
    all of it), the rank renders its band — peer stores and the ordering step stay in the frame — and reads ITS band back
    into pinned host memory: the host ends up with the whole frame, one band per process, 1/N of the traffic per link in both
    directions. Frames are in flight like in e2e_single: the geometry lives in two buffer sets, frame k+1's upload + all-gather
    run on a copy stream while frame k renders, and frame k's band travels to the host on a third stream (a second device
    object) while frame k+1 renders; events order the streams both ways."""
    from cpvulkan_b200.device import Device
    scene, dev, sod, rank, world = rig.work.scene, rig.dev, rig.sod, rig.rank, rig.world
    stream = torch.cuda.current_stream()
    copy_stream = torch.cuda.Stream()
    # the frame's geometry as one byte string (vertex buffers, then the index buffer, each 256-byte aligned), cut into N shards:
    # one pinned-memory copy and one all-gather per frame whatever the number of buffers
    names = [n for n in ("vb", "ib") if n in scene.buffers]
    offsets, total = {}, 0
    for nme in names:
        offsets[nme] = total
        total += (scene.buffers[nme].nbytes + 255) // 256 * 256
    per = ((total + world - 1) // world + 15) // 16 * 16
    blob = np.zeros(per * world, dtype=np.uint8)
    for nme in names:
        blob[offsets[nme]:offsets[nme] + scene.buffers[nme].nbytes] = scene.buffers[nme]
    fulls = [torch.zeros(per * world, dtype=torch.uint8, device="cuda") for _ in range(2)]  # the gathered copies the draws read
    host = torch.from_numpy(blob[rank * per:(rank + 1) * per].copy()).pin_memory()
    ubo = scene.buffers["ubo"]
    ubo_stage = dev.alloc(ubo.nbytes, host_shadow=True)
    dev.shadow(ubo_stage)[:ubo.nbytes] = ubo
    st = sod.m.state
    saved = ({b: st.vertexBuffers[b] for b in scene.vertex_buffers}, st.indexBuffer)
    y0, y1 = rig.band
    band_bytes = (y1 - y0) * scene.color.pitch
    band_addr = sod.m.addr["color"] + y0 * scene.color.pitch
    stream2 = torch.cuda.Stream()
    dev2 = Device(local, stream=stream2.cuda_stream, stats=False)
    slots = []
    for _ in range(2):
        staging = torch.empty(max(band_bytes, 16), dtype=torch.uint8, device="cuda")
        out_dev = dev2.alloc(max(band_bytes, 16), host_shadow=True)  # only its pinned shadow is used, as the read-back target
        slots.append({"staging": staging, "host": dev2.allocs[out_dev][1], "copied": torch.cuda.Event(), "read": torch.cuda.Event(),
                      "geo_ready": torch.cuda.Event(), "geo_free": torch.cuda.Event()})
        slots[-1]["read"].record(stream2)
        slots[-1]["geo_free"].record(stream)

    def prefetch(k):
        """Frame k's geometry: own shard over PCIe, the rest over NVLink, on the copy stream."""
        sl = slots[k % 2]
        copy_stream.wait_event(sl["geo_free"])  # the draw that read this buffer set (frame k - 2) is done
        with torch.cuda.stream(copy_stream):
            mine = fulls[k % 2][rank * per:(rank + 1) * per]
            if "upload" not in skip:
                mine.copy_(host, non_blocking=True)
            if "allgather" not in skip:
                dist.all_gather_into_tensor(fulls[k % 2], mine)
            sl["geo_ready"].record(copy_stream)

    trace = {} if os.environ.get("CPVK_E2E_TRACE") else None  # host seconds per phase, printed by rank 0 (tuning aid)
    skip = os.environ.get("CPVK_E2E_SKIP", "").split(",")          # tuning aid: leave out "upload", "allgather" or "readback" to see what each costs

    def lap(name, t0):
        if trace is not None:
            trace[name] = trace.get(name, 0.0) + time.perf_counter() - t0
        return time.perf_counter()

    def render(k):
        t = time.perf_counter()
        sl = slots[k % 2]
        stream.wait_event(sl["geo_ready"])
        for b, nme in scene.vertex_buffers.items():
            st.vertexBuffers[b] = fulls[k % 2].data_ptr() + offsets[nme]
        if scene.index_buffer:
            st.indexBuffer = fulls[k % 2].data_ptr() + offsets[scene.index_buffer]
        dev.flush()  # the index and vertex bytes were rewritten behind the library's back (copy_ + NCCL): drop what it remembers of them
        t = lap("flush before", t)
        dev.upload_async(sod.m.addr["ubo"], dev.allocs[ubo_stage][1], ubo.nbytes)
        rig.frame()
        t = lap("frame", t)
        dev.flush()                                             # events of this script go behind the draw: have it validated first
        t = lap("flush behind", t)
        sl["geo_free"].record(stream)
        if band_bytes and "readback" not in skip:
            stream.wait_event(sl["read"])                       # the staging buffer's previous contents have reached the host
            dev.copy_rows(sl["staging"].data_ptr(), band_bytes, band_addr, band_bytes, band_bytes, 1)
            sl["copied"].record(stream)
            stream2.wait_event(sl["copied"])
            dev2.download_into_async(sl["host"], sl["staging"].data_ptr(), band_bytes)
            sl["read"].record(stream2)
        lap("read-back", t)

    def run(count):
        prefetch(0)
        for k in range(count):
            if k + 1 < count:
                t = time.perf_counter()
                prefetch(k + 1)  # every rank issues its collectives in this same order
                lap("prefetch", t)
            render(k)

    run(4)
    torch.cuda.synchronize()
    if band_bytes and "readback" not in skip:
        want = dev.download(band_addr, band_bytes)
        for sl in slots:
            got = np.ctypeslib.as_array(C.cast(sl["host"], C.POINTER(C.c_uint8)), shape=(band_bytes,))
            if not np.array_equal(got, want):
                raise SystemExit("e2e read-back of rank %d's band differs from the resident frame" % rank)
    n = max(steps, 200) if rig.work.key == "c3" else steps
    dist.barrier(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    run(n)
    dist.barrier(); torch.cuda.synchronize()  # all streams have drained
    if trace is not None and rank == 0:
        print("e2e host ms per frame by phase: " + ", ".join("%s %.3f" % (k, v * 1e3 / (n + 4)) for k, v in trace.items()), file=sys.stderr)
    t = torch.tensor([(time.perf_counter() - t0) * 1e3 / n], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    for b in scene.vertex_buffers:
        st.vertexBuffers[b] = saved[0][b]
    st.indexBuffer = saved[1]
    dev.flush()
    dev2.close()
    h2d = sum(scene.buffers[nme].nbytes for nme in names) + ubo.nbytes * world
    return {"ms_per_step": float(t[0]), "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": scene.color.nbytes, "frames_timed": n,
            "through": "C ABI on every rank: 1/N of the geometry per rank from pinned host memory + NCCL all-gather over NVLink (copy stream, "
                       "two buffer sets: frame k+1's geometry arrives while frame k renders), clear / draw with the fused peer stores + the ordering "
                       "step, copy_rows of the rank's band to a staging buffer, mem_download_async on a second device object; max over ranks"}


def icd_leg(scene, frames=40):
    """ms/frame as BASELINE.md §4 defines it: wall clock vkQueueSubmit -> fence signalled, through the Vulkan ICD
    (libCPVulkan_b200.so loaded by the loader harness via its manifest; reference: Queue.cpp:11-77). The command buffer holds
    render pass (clear + vkCmdDrawIndexed) + vkCmdCopyImageToBuffer, so the frame is in host-visible memory at the fence; the
    harness also rewrites its vertex data through the mapping every frame, like an application would."""
    import tempfile
    try:
        with tempfile.TemporaryDirectory() as tmp:
            _, _, info = scenes.run_icd(scene, tmp, frames=frames)
        return {"ms_per_frame_icd": info["ms_submit_to_fence"], "ms_per_frame_icd_with_vertex_rewrite": info["ms_per_frame"], "frames": frames,
                "what": "cpvk_harness -> libCPVulkan_b200.so: median wall clock of vkQueueSubmit .. vkWaitForFences (render pass + copy to the host-visible read-back buffer)"}
    except Exception as e:  # never lose the bench line over the extra leg
        return {"error": "%s: %s" % (type(e).__name__, e)}


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

    build.build_cuda()
    # One explicit stream for everything: the draw kernels, the NCCL collectives and the timing events. (torch's default
    # stream has handle 0, which the C ABI reads as "use the device's own stream" — events recorded on the default
    # stream would then not be ordered against the kernels.)
    stream = torch.cuda.Stream()
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0

    work = Workload(args.config)
    rig = Rig(work, torch, dist, stream, local, rank, world, args.gather, args.barrier)
    if world > 1:
        verify_exchange(rig)
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    m = measure(rig, args.steps, args.warmup, sampler)
    clocks = sampler.stop() if sampler else None
    ms_step, ms_raster = m["ms_step"], m["kernel_ms"]["raster"]
    b_alg = work.algorithmic_bytes(m["local_cov"], m["local_pass"])
    peak, peak_src = measured_peaks()
    achieved = b_alg / (ms_raster * 1e-3) / 1e9 if ms_raster > 0 else 0.0

    e2e = e2e_multi(torch, dist, rig, local, args.steps) if world > 1 else e2e_single(torch, rig, local, args.steps)
    e2e["value"] = work.units(m["n_cov"]) / (e2e["ms_per_step"] * 1e-3) / work.scale
    e2e["unit"] = work.unit

    # the other BASELINE configs, short, on every N (SCALE_rNN.json then holds C4's scaling next to C3's)
    extras = None
    if not args.no_extras:
        extras = secondary_configs(args, torch, dist, stream, local, rank, world, rig)
    line = None
    if rank == 0:
        cpu = None
        if world == 1 and not args.no_cpu:
            build.build_oracle()
            threads = max(1, len(os.sched_getaffinity(0)))
            scene, factor, what = cpu_sample(work, threads)
            t_cpu, cov_cpu = oracle_draw_seconds(scene, threads, repeats=3 if work.key == "c3" else 1)
            t_one, _ = oracle_draw_seconds(scene, 1)
            model, ncpu = cpu_info()
            cpu = {"value": work.units(int(cov_cpu * factor)) / (t_cpu * factor) / work.scale, "unit": work.unit, "cores": threads, "kind": "port",
                   "sample": "%s, median of runs, %d threads (one horizontal window each); one thread: %.3f %s; %s, %d logical CPUs; optimistic stand-in "
                             "for the reference ICD (no JIT / indirection overhead)" % (what, threads, work.units(int(cov_cpu * factor)) / (t_one * factor) / work.scale, work.unit, model, ncpu),
                   "fragments_match_gpu": int(cov_cpu * factor) == m["n_cov"]}
        icd = icd_leg(work.scene) if (world == 1 and work.key == "c3" and not args.no_extras) else None
        exchange = ""
        if world > 1:
            exchange = " + bands stored by k_raster over NVLink into %s + %s" % ("every GPU's frame" if args.gather == "all" else "the presenting GPU's frame (rank 0)",
                                                                                     "cpvk_cuda_peer_barrier" if args.barrier == "flags" else "4-byte NCCL all-reduce")
        blocks = m["blocks"]
        line = {
            "metric": work.metric, "value": work.units(m["n_cov"]) / (ms_step * 1e-3) / work.scale, "unit": work.unit,
            "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": work.workload + exchange, "parallelism": "sort-first bands of tile rows x%d" % world, "l2": work.l2, "cpu_affinity": affinity},
            "blocks": {"count": len(blocks), "frames_timed": len(blocks) * args.steps, "seconds_timed": sum(blocks) / 1e3,
                       "ms_per_step_min": min(blocks) / args.steps, "ms_per_step_median": ms_step, "ms_per_step_max": max(blocks) / args.steps},
            "mtris_per_s": work.prims / (ms_step * 1e-3) / 1e6, "gfragments_per_s": m["n_cov"] / (ms_step * 1e-3) / 1e9, "ms_per_frame": ms_step,
            "fragments_covered": m["n_cov"], "fragments_written": m["n_pass"], "bin_entries_rank0": m["bin_entries"], "host_issue_ms_rank0": m["host_issue_ms"],
            "kernel_ms_rank0": m["kernel_ms"],
            "roofline": {"bound": "hbm", "kernel": "cpvk_k_raster", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "peak_source": peak_src, "algorithmic_bytes_per_launch": b_alg, "traffic": raster_traffic(work, world),
                         "note": "rank 0's launch; C3/M1 (8 px per triangle) is instruction-bound, C4 fragment-stage bound; see DESIGN.md §6"},
            "cpu_baseline": cpu, "e2e": e2e, "clocks": clocks, "gpu_launches": int(round(m["launches_per_step"] * args.steps)),
            "gpu_launches_per_step": m["launches_per_step"],
        }
        if icd:
            line["icd"] = icd
            if "ms_per_frame_icd" in icd:
                line["ms_per_frame_icd"] = icd["ms_per_frame_icd"]
        if extras:
            line["other_configs"] = extras
    rig.close()
    if world > 1:
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), file=_RESULT_OUT, flush=True)


def raster_traffic(work, world):
    """DRAM bytes of one cpvk_k_raster launch from the committed `ncu --set full` capture of this very command — only when that
    capture was taken on the same workload and GPU count; never measured inside a bench run, null otherwise."""
    try:
        with open(os.path.join(ROOT, "profiles", "raster_traffic.json")) as f:
            t = json.load(f)
        if t.get("workload") != work.key or int(t.get("n_gpus", 1)) != world:
            return None
        return t["dram_bytes_read"] + t["dram_bytes_write"]
    except (OSError, KeyError, ValueError):
        return None


def secondary_configs(args, torch, dist, stream, local, rank, world, main_rig):
    """Short measurements of the other BASELINE configs: C4 (a slice of its 2,000 quads, the whole 8K frame) on EVERY N with the
    same band split and exchange as the headline; C5 (vkCmdBlitImage / vkCmdCopyImage at 7680x4320) on one GPU. CUDA events on
    the launching stream, inputs resident. All ranks take part; rank 0 returns the dict."""
    peak, _ = measured_peaks()
    out = {}
    if args.config == "c3":
        quads = 100
        w4 = Workload("c4", quads=quads)
        rig = Rig(w4, torch, dist, stream, local, rank, world, args.gather, args.barrier)
        if world > 1:
            verify_exchange(rig)
        m = measure_short(rig, reps=3)
        rig.close()
        out["C4_overdraw"] = {"workload": w4.workload + " (%d of the 2,000 quads; cost is linear in quads)" % quads, "n_gpus": world,
                              "ms_per_step": m["ms"], "gfragments_per_s": m["n_cov"] / (m["ms"] * 1e-3) / 1e9, "mtris_per_s": 2 * quads / (m["ms"] * 1e-3) / 1e6,
                              "ms_per_frame_at_2000_quads": m["ms"] * 2000 / quads,
                              "roofline_frac_whole_job": m["n_cov"] * 16 / (m["ms"] * 1e-3) / 1e9 / (peak * world), "algorithmic_bytes_per_fragment": 16}
    if world > 1 or rank != 0:
        return out if rank == 0 else None
    dev = main_rig.dev

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

    W, H = 7680, 4320

    def image(fmt, w, h, texel):
        t = torch.randint(0, 255, (w * h * texel,), dtype=torch.uint8, device="cuda")
        return t, capi.Attachment(t.data_ptr(), w, h, w * texel, fmt)

    s8, a8 = image(37, W, H, 4)
    d16, a16 = image(97, W, H, 8)
    s4, a4 = image(37, W // 2, H // 2, 4)
    s4h, a4h = image(97, W // 2, H // 2, 8)
    d8, _ = image(37, W, H, 4)
    b1 = capi.Blit(a8, a16, 0, 0, W, H, 0, 0, W, H, 0)
    b2 = capi.Blit(a4, a16, 0, 0, W // 2, H // 2, 0, 0, W, H, 1)
    b3 = capi.Blit(a4h, a16, 0, 0, W // 2, H // 2, 0, 0, W, H, 1)
    if args.config == "c3":
        # the headline draw with two frames in flight (two device objects, own streams and targets, alternating): the idle tails
        # between one frame's kernels are filled by the other frame's — what a double-buffered application gets, wall clock
        from cpvulkan_b200.device import Device, SceneOnDevice
        lanes = []
        for _ in range(2):
            ldev = Device(local, stats=False)
            lanes.append((ldev, SceneOnDevice(ldev, main_rig.work.scene)))
        def frames(count):
            for k in range(count):
                lanes[k % 2][1].clear(); lanes[k % 2][1].draw()
            for ldev, _ in lanes:
                ldev.sync()
        frames(8)
        t0 = time.perf_counter(); frames(400); ms = (time.perf_counter() - t0) * 1e3 / 400
        for ldev, lsod in lanes:
            lsod.close(); ldev.close()
        out["C3_two_frames_in_flight"] = {"ms_per_frame": ms, "mtris_per_s": main_rig.work.prims / (ms * 1e-3) / 1e6, "frames_timed": 400,
                                          "note": "resident inputs, two device objects alternate frames on their own streams; wall clock"}
    for key, fn, nbytes in (("blit_8k_rgba8_to_rgba16f_nearest", lambda: dev.blit(b1), W * H * 12),
                            ("blit_4k_rgba8_to_8k_rgba16f_linear", lambda: dev.blit(b2), W * H * 8 + W * H),
                            ("blit_4k_to_8k_rgba16f_linear", lambda: dev.blit(b3), W * H * 8 + W * H * 2),
                            ("copy_image_8k_rgba8", lambda: dev.copy_rows(d8.data_ptr(), W * 4, s8.data_ptr(), W * 4, W * 4, H), W * H * 8)):
        ms = timed(fn, 10)
        out["C5_" + key] = {"ms": ms, "GBps": nbytes / ms / 1e6, "roofline_frac": nbytes / ms / 1e6 / peak, "algorithmic_bytes": nbytes}
    return out


def measure_short(rig, reps):
    torch, dist, world = rig.torch, rig.dist, rig.world
    for _ in range(2):
        rig.frame()
    rig.dev.flush()
    torch.cuda.synchronize()
    st = rig.dev.stats()
    n_cov = int(st.fragmentsCovered)
    if world > 1:
        t = torch.tensor([n_cov], dtype=torch.int64, device="cuda")
        dist.all_reduce(t)
        n_cov = int(t[0])
    rig.dev.set_stats(False)
    blocks = timed_blocks(torch, dist, world, rig.frame, reps, rig.dev.flush, min_seconds=0.0, min_frames=reps, min_blocks=3, max_blocks=3)
    return {"ms": statistics.median(blocks) / reps, "n_cov": n_cov}


_RESULT_OUT = sys.stdout


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c3", choices=["c3", "c4"], help="c3 (default): the 1M-triangle 4K draw; c4: all 2,000 blended quads at 8K")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--barrier", default="flags", choices=["flags", "nccl"], help="multi-GPU per-frame ordering step: cpvk_cuda_peer_barrier (flag words in peer-mapped memory) or a 4-byte NCCL all-reduce")
    ap.add_argument("--gather", default="one", choices=["one", "all"], help="multi-GPU exchange target of k_raster's peer stores: the presenting GPU (rank 0) or every GPU")
    ap.add_argument("--no-extras", action="store_true", help="skip the short C4 / C5 measurements (other_configs) and the ICD leg")
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

"""Where the e2e frame time of C3/M1 goes: the bench's e2e loop (bench.py e2e_single) with parts switched off.
  U = upload the frame's inputs, R = clear + draw, D = read the colour attachment back; lanes = device objects alternating."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpvulkan_b200 import scenes
from cpvulkan_b200.device import Device, SceneOnDevice

scene = scenes.mesh_indexed()
names = [n for n in ("vb", "ib", "ubo") if n in scene.buffers]
out = {}
for n_lanes in (1, 2, 3):
    lanes = []
    for i in range(n_lanes):
        ldev = Device(0, stats=False)
        lsod = SceneOnDevice(ldev, scene)
        staged = {}
        for nme in names:
            data = scene.buffers[nme]
            a = ldev.alloc(data.nbytes, host_shadow=True)
            ldev.shadow(a)[:data.nbytes] = data
            staged[nme] = (a, data.nbytes)
        out_dev = ldev.alloc(scene.color.nbytes, host_shadow=True)
        lanes.append((ldev, lsod, staged, ldev.allocs[out_dev][1]))
    for parts in ("U", "R", "D", "UD", "UR", "RD", "URD"):
        def step(k):
            ldev, lsod, staged, out_host = lanes[k % n_lanes]
            ldev.sync()
            if "U" in parts:
                for nme in names:
                    src_alloc, nbytes = staged[nme]
                    ldev.upload_async(lsod.m.addr[nme], ldev.allocs[src_alloc][1], nbytes)
            if "R" in parts:
                lsod.clear(); lsod.draw()
            if "D" in parts:
                ldev.download_into_async(out_host, lsod.m.addr["color"], scene.color.nbytes)
        for k in range(6):
            step(k)
        for l in lanes:
            l[0].sync()
        n = 100
        t0 = time.perf_counter()
        for k in range(n):
            step(k)
        t_host = (time.perf_counter() - t0) * 1e3 / n
        for l in lanes:
            l[0].sync()
        out["%d lanes %s" % (n_lanes, parts)] = {"ms": round((time.perf_counter() - t0) * 1e3 / n, 4), "host_ms": round(t_host, 4)}
    for l in lanes:
        l[1].close(); l[0].close()
print(json.dumps(out, indent=1))

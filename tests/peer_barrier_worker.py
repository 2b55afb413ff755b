"""One side of tests/test_group_gpu.py's two-process barrier tests: python peer_barrier_worker.py <producer|consumer> <copy|draw> <dir>.
The two processes share GPU 0; handles (cpvk_cuda_mem_export) travel through files in <dir>."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from cpvulkan_b200 import scenes
from cpvulkan_b200.device import Device, SceneOnDevice

role, scenario, folder = sys.argv[1], sys.argv[2], sys.argv[3]
me = 0 if role == "producer" else 1
SIZE = 1 << 20
ROUNDS = 8 if scenario == "copy" else 3


def publish(name, handle):
    tmp = os.path.join(folder, name + ".tmp")
    with open(tmp, "wb") as f:
        f.write(handle)
    os.rename(tmp, os.path.join(folder, name))


def fetch(name, timeout=120.0):
    path = os.path.join(folder, name)
    t0 = time.time()
    while not os.path.exists(path):
        if time.time() - t0 > timeout:
            raise SystemExit("%s: %s never appeared" % (role, name))
        time.sleep(0.02)
    with open(path, "rb") as f:
        return f.read()


dev = Device(0, stats=False)
flags = dev.alloc(64)
dev.sync()  # zero-filled before the peer can signal into it
publish("flags_%d" % me, dev.export_handle(flags))
other = dev.import_handle(fetch("flags_%d" % (1 - me)))
arrays = [flags, other] if me == 0 else [other, flags]

if scenario == "copy":
    if role == "producer":
        src, data = dev.alloc(SIZE), dev.alloc(SIZE)
        dev.sync()
        publish("data", dev.export_handle(data))
        for k in range(ROUNDS):
            dev.upload(src, np.random.default_rng(k).integers(0, 256, SIZE, dtype=np.uint8))
            dev.copy_rows(data, SIZE, src, SIZE, SIZE, 1)
            dev.peer_barrier(arrays, me, 2 * k + 1)     # the consumer may read `data`
            dev.peer_barrier(arrays, me, 2 * k + 2)     # ... and is done with it
        dev.sync()
        print("producer ok")
    else:
        out = dev.alloc(SIZE)
        data = dev.import_handle(fetch("data"))
        for k in range(ROUNDS):
            dev.peer_barrier(arrays, me, 2 * k + 1)
            dev.copy_rows(out, SIZE, data, SIZE, SIZE, 1)
            dev.peer_barrier(arrays, me, 2 * k + 2)
            got = dev.download(out, SIZE)
            want = np.random.default_rng(k).integers(0, 256, SIZE, dtype=np.uint8)
            if not np.array_equal(got, want):
                raise SystemExit("round %d: the consumer ran ahead of the producer" % k)
        dev.unimport(data)
        print("consumer ok")
else:
    scene = scenes.random_triangles(width=64, height=64, tris=900, seed=51)
    nbytes = scene.color.nbytes
    if role == "producer":
        sod = SceneOnDevice(dev, scene)
        dev.sync()
        publish("data", dev.export_handle(sod.m.addr["color"]))
        for k in range(ROUNDS):
            sod.clear(); sod.draw()
            dev.peer_barrier(arrays, me, 2 * k + 1)     # behind the unsettled draw: round 0's is replayed
            dev.peer_barrier(arrays, me, 2 * k + 2)
            dev.sync()
        sod.close()
        print("producer ok")
    else:
        want, _, _ = scenes.run_oracle(scene)
        out = dev.alloc(nbytes)
        data = dev.import_handle(fetch("data"))
        for k in range(ROUNDS):
            dev.peer_barrier(arrays, me, 2 * k + 1)
            dev.copy_rows(out, nbytes, data, nbytes, nbytes, 1)
            dev.peer_barrier(arrays, me, 2 * k + 2)
            if not np.array_equal(dev.download(out, nbytes), want):
                raise SystemExit("round %d: the barrier let the consumer through before the frame was complete" % k)
        dev.unimport(data)
        print("consumer ok")
dev.unimport(other)
dev.close()

"""Per-kernel device times of the C3/M1 draw (library timing mode), for A/B runs of front-end variants."""
import json, os, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpvulkan_b200 import scenes
from cpvulkan_b200.device import Device, SceneOnDevice

dev = Device(0, stats=False, timing=True)
sod = SceneOnDevice(dev, scenes.mesh_indexed())
acc = {"vertex": [], "setup": [], "bin": [], "raster": []}
for i in range(30):
    sod.clear(); sod.draw()
    st = dev.stats()
    if i >= 5:
        acc["vertex"].append(st.msVertex); acc["setup"].append(st.msSetup); acc["bin"].append(st.msBin); acc["raster"].append(st.msRaster)
print(json.dumps({k: round(statistics.median(v) * 1e3, 1) for k, v in acc.items()}))
sod.close(); dev.close()

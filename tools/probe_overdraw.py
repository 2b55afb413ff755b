import sys, time, json
sys.path.insert(0, __import__("os").path.dirname(__import__("os").path.dirname(__import__("os").path.abspath(__file__))))
import torch
from cpvulkan_b200 import scenes
from cpvulkan_b200.device import Device, SceneOnDevice
dev = Device(0, stats=True, timing=True)
for quads, w, h in ((20, 7680, 4320), (200, 7680, 4320)):
    sc = scenes.overdraw_quads(width=w, height=h, quads=quads, tex_size=1024)
    t0 = time.time(); s = SceneOnDevice(dev, sc); t1 = time.time()
    s.render(); dev.sync()
    s.clear(); s.draw(); st = dev.stats()
    frags = st.fragmentsCovered
    print(json.dumps({"quads": quads, "pipeline_s": round(t1 - t0, 2), "frags": frags, "ms": {k: round(getattr(st, k), 3) for k in ("msVertex", "msSetup", "msBin", "msRaster", "msTotal")},
                      "gfrag_s": round(frags / (st.msTotal * 1e-3) / 1e9, 2), "bin_entries": st.binEntries,
                      "roofline_frac_16B": round(frags * 16 / (st.msRaster * 1e-3) / 6542.1e9, 4)}))
    s.close()
dev.close()

"""One blended, linearly-filtered textured overdraw draw (C4 shape at reduced size) for ncu captures of cpvk_k_raster."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpvulkan_b200 import scenes  # noqa: E402
from cpvulkan_b200.device import Device, SceneOnDevice  # noqa: E402

dev = Device(0, stats=True, timing=True)
s = SceneOnDevice(dev, scenes.overdraw_quads(width=3840, height=2160, quads=10, tex_size=1024))
for _ in range(3):
    s.render()
dev.sync()
st = dev.stats()
print("frags", st.fragmentsCovered, "raster ms", st.msRaster, "Gfrag/s", st.fragmentsCovered / (st.msRaster * 1e-3) / 1e9)
s.close()
dev.close()

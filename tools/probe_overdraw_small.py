"""One small C4-shaped draw (20 blended, LINEAR-filtered full-screen quads at 4K RGBA16F) for ncu captures."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from cpvulkan_b200 import scenes
from cpvulkan_b200.device import Device, SceneOnDevice
dev = Device(0, stats=False, timing=False)
sc = scenes.overdraw_quads(width=3840, height=2160, quads=10, tex_size=1024)
s = SceneOnDevice(dev, sc)
for _ in range(3):
    s.render()
dev.sync()
s.close(); dev.close()

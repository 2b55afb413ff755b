"""C5 at full size on the GPU: vkCmdBlitImage / vkCmdCopyImage / clear through the C ABI.
Times with CUDA events on the device's stream and reports achieved GB/s against the algorithmic bytes
(source bytes read once + destination bytes written once)."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from cpvulkan_b200 import capi
from cpvulkan_b200.device import Device

RGBA8, RGBA16F = 37, 97
stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
dev = Device(0, stream=stream.cuda_stream, stats=False)
dev.set_lazy_clear(False)


def image(fmt, w, h, texel):
    t = torch.randint(0, 255, (w * h * texel,), dtype=torch.uint8, device="cuda")
    return t, capi.Attachment(t.data_ptr(), w, h, w * texel, fmt)


def timed(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


W, H = 7680, 4320
out = {}
s8, a8 = image(RGBA8, W, H, 4)
d16, a16 = image(RGBA16F, W, H, 8)
b = capi.Blit(a8, a16, 0, 0, W, H, 0, 0, W, H, 0)
ms = timed(lambda: dev.blit(b)); out["blit_8k_rgba8_to_rgba16f_nearest"] = {"ms": ms, "GBps": (W * H * 12) / ms / 1e6}
s4, a4 = image(RGBA8, W // 2, H // 2, 4)
b2 = capi.Blit(a4, a16, 0, 0, W // 2, H // 2, 0, 0, W, H, 1)
ms = timed(lambda: dev.blit(b2)); out["blit_4k_to_8k_rgba16f_linear"] = {"ms": ms, "GBps": (W * H * 8 + W * H) / ms / 1e6}
d8, ad8 = image(RGBA8, W, H, 4)
ms = timed(lambda: dev.copy_rows(d8.data_ptr(), W * 4, s8.data_ptr(), W * 4, W * 4, H)); out["copy_8k_rgba8"] = {"ms": ms, "GBps": (W * H * 8) / ms / 1e6}
cv = capi.ClearValue()
ms = timed(lambda: dev.clear(a16, cv, 0)); out["clear_8k_rgba16f"] = {"ms": ms, "GBps": (W * H * 8) / ms / 1e6}
print(json.dumps(out))
dev.close()

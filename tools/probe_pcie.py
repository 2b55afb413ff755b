"""What the host link gives on this box: pinned host <-> HBM copies of the e2e leg's sizes (28 MB up, 33 MB down), each
direction alone and both at once on two streams. The e2e floor of C3/M1 is max(up, down) per frame when the two overlap."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

up_n, down_n = 28048096, 33177600
hu = torch.empty(up_n, dtype=torch.uint8).pin_memory(); du = torch.empty(up_n, dtype=torch.uint8, device="cuda")
hd = torch.empty(down_n, dtype=torch.uint8).pin_memory(); dd = torch.empty(down_n, dtype=torch.uint8, device="cuda")
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def up():
    with torch.cuda.stream(s1):
        du.copy_(hu, non_blocking=True)


def down():
    with torch.cuda.stream(s2):
        hd.copy_(dd, non_blocking=True)


def both():
    up(); down()


out = {}
ms = timed(up); out["h2d"] = {"ms": ms, "GBps": up_n / ms / 1e6}
ms = timed(down); out["d2h"] = {"ms": ms, "GBps": down_n / ms / 1e6}
ms = timed(both); out["both"] = {"ms": ms, "GBps_up": up_n / ms / 1e6, "GBps_down": down_n / ms / 1e6}
print(json.dumps(out))

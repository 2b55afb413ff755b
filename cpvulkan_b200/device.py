"""Thin Python driver over the C ABI (libcpvk_cuda.so) — what tests, smoke() and bench.py call.

No fallback: constructing a Device without the built library or without a CUDA device raises.
"""
import ctypes as C

import numpy as np

from . import capi, scenes


class CpvkError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("cpvk_cuda error %d: %s" % (code, msg))
        self.code = code


def _check(lib, rc):
    if rc != 0:
        raise CpvkError(rc, (lib.cpvk_cuda_last_error() or b"").decode(errors="replace"))


class Device:
    def __init__(self, ordinal=0, stream=None, stats=True, timing=False, group=None):
        """group = list of CUDA ordinals: one handle driving several GPUs of the box (cpvk_cuda_device_create_group)."""
        self.lib = capi.load_cuda()
        self.handle = C.c_void_p()
        if group is not None:
            ords = (C.c_int * len(group))(*group)
            _check(self.lib, self.lib.cpvk_cuda_device_create_group(ords, len(group), C.byref(self.handle)))
        else:
            _check(self.lib, self.lib.cpvk_cuda_device_create(ordinal, C.byref(self.handle)))
        if stream is not None:
            _check(self.lib, self.lib.cpvk_cuda_device_set_stream(self.handle, C.c_void_p(stream)))
        _check(self.lib, self.lib.cpvk_cuda_device_set_stats(self.handle, int(stats)))
        _check(self.lib, self.lib.cpvk_cuda_device_set_timing(self.handle, int(timing)))
        self.allocs = {}

    def close(self):
        if self.handle:
            self.lib.cpvk_cuda_device_destroy(self.handle)
            self.handle = None

    def set_timing(self, on):
        _check(self.lib, self.lib.cpvk_cuda_device_set_timing(self.handle, int(on)))

    def set_stats(self, on):
        _check(self.lib, self.lib.cpvk_cuda_device_set_stats(self.handle, int(on)))

    def set_lazy_clear(self, on):
        """Deferred clears (default on): see cpvk_cuda_flush in include/cpvk_cuda.h."""
        _check(self.lib, self.lib.cpvk_cuda_device_set_lazy_clear(self.handle, int(on)))

    def set_speculation(self, on):
        """Speculative draw tails (default on): see cpvk_cuda_device_set_speculation in include/cpvk_cuda.h."""
        _check(self.lib, self.lib.cpvk_cuda_device_set_speculation(self.handle, int(on)))

    def set_overlap(self, on):
        """Front-end overlap across draws (default: on for the device's own stream): see cpvk_cuda_device_set_overlap."""
        _check(self.lib, self.lib.cpvk_cuda_device_set_overlap(self.handle, int(on)))

    def flush(self):
        _check(self.lib, self.lib.cpvk_cuda_flush(self.handle))

    # memory ------------------------------------------------------------------------------------
    def alloc(self, nbytes, host_shadow=False):
        dev = C.c_uint64()
        host = C.c_void_p()
        _check(self.lib, self.lib.cpvk_cuda_mem_alloc(self.handle, max(nbytes, 1), C.byref(dev), C.byref(host) if host_shadow else None))
        self.allocs[dev.value] = (nbytes, host.value)
        return dev.value

    def shadow(self, addr):
        """numpy view of the pinned host shadow of an allocation made with host_shadow=True."""
        nbytes, host = self.allocs[addr]
        return np.ctypeslib.as_array(C.cast(host, C.POINTER(C.c_uint8)), shape=(max(nbytes, 1),))

    def free(self, addr):
        _check(self.lib, self.lib.cpvk_cuda_mem_free(self.handle, addr))
        self.allocs.pop(addr, None)

    def upload(self, addr, data):
        data = np.ascontiguousarray(data).view(np.uint8).reshape(-1)
        _check(self.lib, self.lib.cpvk_cuda_mem_upload(self.handle, addr, data.ctypes.data_as(C.c_void_p), data.nbytes))
        self.sync()  # `data` may be pageable and short-lived

    def upload_async(self, addr, host_ptr, nbytes):
        _check(self.lib, self.lib.cpvk_cuda_mem_upload(self.handle, addr, C.c_void_p(host_ptr), nbytes))

    def download(self, addr, nbytes):
        out = np.empty(max(nbytes, 1), dtype=np.uint8)
        _check(self.lib, self.lib.cpvk_cuda_mem_download(self.handle, out.ctypes.data_as(C.c_void_p), addr, nbytes))
        return out[:nbytes]

    def download_into(self, host_ptr, addr, nbytes):
        _check(self.lib, self.lib.cpvk_cuda_mem_download(self.handle, C.c_void_p(host_ptr), addr, nbytes))

    def download_into_async(self, host_ptr, addr, nbytes):
        _check(self.lib, self.lib.cpvk_cuda_mem_download_async(self.handle, C.c_void_p(host_ptr), addr, nbytes))

    def sync(self):
        _check(self.lib, self.lib.cpvk_cuda_sync(self.handle))

    # pipeline / commands -------------------------------------------------------------------------
    def create_pipeline(self, desc):
        p = C.c_void_p()
        _check(self.lib, self.lib.cpvk_cuda_pipeline_create(self.handle, C.byref(desc), C.byref(p)))
        return p

    def destroy_pipeline(self, p):
        self.lib.cpvk_cuda_pipeline_destroy(self.handle, p)

    def draw(self, state):
        _check(self.lib, self.lib.cpvk_cuda_draw(self.handle, C.byref(state)))

    def stats(self):
        st = capi.DrawStats()
        _check(self.lib, self.lib.cpvk_cuda_last_draw_stats(self.handle, C.byref(st)))
        return st

    def clear(self, attachment, value, is_depth_stencil):
        _check(self.lib, self.lib.cpvk_cuda_clear(self.handle, C.byref(attachment), C.byref(value), int(is_depth_stencil)))

    def copy_rows(self, dst, dst_pitch, src, src_pitch, row_bytes, rows):
        _check(self.lib, self.lib.cpvk_cuda_copy_rows(self.handle, dst, dst_pitch, src, src_pitch, row_bytes, rows))

    def blit(self, blit):
        _check(self.lib, self.lib.cpvk_cuda_blit(self.handle, C.byref(blit)))

    def gather(self, attachments):
        """After the draws of a pass on a group: exchange the members' bands of these images (no-op on a plain device)."""
        arr = (capi.Attachment * len(attachments))(*attachments)
        _check(self.lib, self.lib.cpvk_cuda_gather(self.handle, arr, len(attachments)))

    def group_size(self):
        return int(self.lib.cpvk_cuda_group_size(self.handle))

    def export_handle(self, addr):
        h = (C.c_uint8 * 64)()
        _check(self.lib, self.lib.cpvk_cuda_mem_export(self.handle, addr, h))
        return bytes(h)

    def import_handle(self, handle):
        out = C.c_uint64()
        buf = (C.c_uint8 * 64).from_buffer_copy(handle)
        _check(self.lib, self.lib.cpvk_cuda_mem_import(self.handle, buf, C.byref(out)))
        return out.value

    def unimport(self, addr):
        _check(self.lib, self.lib.cpvk_cuda_mem_unimport(self.handle, addr))

    def peer_barrier(self, flag_arrays, self_index, sequence):
        """cpvk_cuda_peer_barrier: device-side ordering between the GPUs of a one-process-per-GPU run."""
        key = tuple(flag_arrays)
        if getattr(self, "_barrier_key", None) != key:  # the same participants frame after frame: build the argument once
            self._barrier_key, self._barrier_arr = key, (C.c_uint64 * len(flag_arrays))(*flag_arrays)
        _check(self.lib, self.lib.cpvk_cuda_peer_barrier(self.handle, self._barrier_arr, len(flag_arrays), self_index, sequence))

    def launch_count(self):
        return int(self.lib.cpvk_cuda_launch_count(self.handle))


class SceneOnDevice:
    """A scenes.Scene resident in HBM: buffers uploaded once, pipeline linked once, re-drawable."""

    def __init__(self, dev, scene, band=None):
        self.dev, self.scene = dev, scene
        self.owned = []

        def alloc(name, nbytes, init):
            a = dev.alloc(nbytes)
            self.owned.append(a)
            if init is not None:
                dev.upload(a, np.ascontiguousarray(init).view(np.uint8).reshape(-1)[:nbytes])
            return a

        self.m = scenes.materialize(scene, alloc)
        self.pipeline = dev.create_pipeline(self.m.desc)
        self.m.state.pipeline = self.pipeline.value
        if band is not None:
            self.m.state.bandY0, self.m.state.bandY1 = band

    def clear(self, band_only=False):
        """Render-pass clear of the attachments. band_only: just the rows of this GPU's sort-first band (the other
        rows belong to other GPUs and are overwritten by the gather)."""
        y0, y1 = self.m.state.bandY0, self.m.state.bandY1
        key = (band_only, y0, y1, getattr(self.m.color_attachment, "address", 0), getattr(self.m.depth_attachment, "address", 0))
        if getattr(self, "_clear_key", None) != key:  # a frame loop clears the same rectangles every frame: build the arguments once
            self._clear_key, self._clear_args = key, []
            for img, att in ((self.scene.color, self.m.color_attachment), (self.scene.depth, self.m.depth_attachment)):
                if img is not None and img.clear is not None:
                    cv, is_ds = scenes.clear_value(img)
                    if band_only and y1 > y0:
                        att = capi.Attachment(att.address + y0 * att.rowPitch, att.width, min(y1, att.height) - y0, att.rowPitch, att.format)
                    self._clear_args.append((att, cv, is_ds))
        for att, cv, is_ds in self._clear_args:
            self.dev.clear(att, cv, is_ds)

    def draw(self):
        self.dev.draw(self.m.state)

    def render(self):
        self.clear()
        self.draw()
        self.gather()

    def gather(self):
        """Group devices: bring every member's band of the attachments to every replica (no-op on one GPU)."""
        atts = [a for img, a in ((self.scene.color, self.m.color_attachment), (self.scene.depth, self.m.depth_attachment)) if img is not None]
        self.dev.gather(atts)

    def read_color(self):
        return self.dev.download(self.m.addr["color"], self.scene.color.nbytes)

    def read_depth(self):
        return self.dev.download(self.m.addr["depth"], self.scene.depth.nbytes) if self.scene.depth else None

    def close(self):
        self.dev.sync()
        self.dev.destroy_pipeline(self.pipeline)
        for a in self.owned:
            self.dev.free(a)
        self.owned = []


def run_cuda(dev, scene, band=None):
    """Render `scene` once on the GPU. Returns (color bytes, depth bytes or None, DrawStats)."""
    s = SceneOnDevice(dev, scene, band)
    try:
        s.render()
        st = dev.stats()
        return s.read_color(), s.read_depth(), st
    finally:
        s.close()

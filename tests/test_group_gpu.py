"""Multi-GPU inside the product: one CpvkDevice handle / one VkDevice drives several GPUs — sort-first bands of
tile rows, replicated resources, bands exchanged by cpvk_cuda_gather (peer copies over NVLink) — byte for byte against the CPU
oracle, through the C ABI and through the Vulkan ICD (CPVK_CUDA_DEVICES). On a box with one GPU the group is built from the
same ordinal listed several times: every member still has its own replicas, stream and band, so address translation, the
band split, the exchange and its event ordering all run; with >= 2 GPUs the same tests run across real peers as well."""
import numpy as np
import pytest

from cpvulkan_b200 import capi, scenes
from cpvulkan_b200.device import Device, SceneOnDevice

pytestmark = pytest.mark.gpu


def gpu_count():
    import torch
    return torch.cuda.device_count()


def groups():
    out = [[0, 0], [0, 0, 0]]
    n = gpu_count()
    if n >= 2:
        out.append(list(range(min(n, 8))))
    return out


@pytest.fixture(scope="module", params=["0,0", "0,0,0", "all"])
def group(request, built):
    if request.param == "all":
        n = gpu_count()
        if n < 2:
            pytest.skip("one GPU visible: the multi-ordinal group runs under gpurun --gpus 2/4/8")
        ords = list(range(min(n, 8)))
    else:
        ords = [int(v) for v in request.param.split(",")]
    d = Device(group=ords, stats=True)
    assert d.group_size() == len(ords)
    yield d
    d.close()


def render_and_compare(dev, scene):
    oc, od, ost = scenes.run_oracle(scene)
    s = SceneOnDevice(dev, scene)
    try:
        s.render()  # clear + draw + gather
        st = dev.stats()
        gc, gd = s.read_color(), s.read_depth()
    finally:
        s.close()
    assert (st.primitives, st.fragmentsCovered, st.fragmentsWritten) == (ost.primitives, ost.fragmentsCovered, ost.fragmentsWritten)
    assert np.array_equal(gc, oc), "colour differs from the oracle (%d bytes)" % int((gc != oc).sum())
    if od is not None:
        assert np.array_equal(gd, od), "depth differs from the oracle"


def test_group_draws_match_the_oracle(group):
    render_and_compare(group, scenes.draw_cube(200, 136))                       # 4.25 tile rows: uneven bands, a partial last tile row
    render_and_compare(group, scenes.draw_textured_cube(160, 100, filt=scenes.LINEAR))
    render_and_compare(group, scenes.random_triangles(width=256, height=192, tris=300, seed=3))
    render_and_compare(group, scenes.overdraw_quads(96, 80, quads=6, tex_size=32))
    render_and_compare(group, scenes.mesh_indexed(640, 360, 160, 90))
    render_and_compare(group, scenes.random_triangles(width=64, height=20, tris=40, seed=4))  # fewer tile rows than members of the largest group
    render_and_compare(group, scenes.random_points_lines(topology=scenes.LINE_LIST))


def test_group_frames_with_changing_content(group):
    """Two different frames in a row into the same attachments: stale rows at band seams (the own-frame write-back
    of a tile straddling a band edge) would show up as bytes of the previous frame."""
    a = scenes.random_triangles(width=200, height=150, tris=250, seed=21)
    b = scenes.random_triangles(width=200, height=150, tris=250, seed=22)
    want = {id(x): scenes.run_oracle(x) for x in (a, b)}
    sa = SceneOnDevice(group, a)
    try:
        for frame in range(4):
            cur = a if frame % 2 == 0 else b
            # same attachments, other geometry: upload the other scene's vertex data into the resident buffers
            for name, data in cur.buffers.items():
                group.upload(sa.m.addr[name], data)
            sa.render()
            assert np.array_equal(sa.read_color(), want[id(cur)][0]), "frame %d" % frame
            assert np.array_equal(sa.read_depth(), want[id(cur)][1]), "frame %d (depth)" % frame
    finally:
        sa.close()


def test_group_two_draws_one_gather(group):
    """Several draws of a pass, one exchange at its end (what the ICD does at vkCmdEndRenderPass)."""
    scene = scenes.random_triangles(width=160, height=120, tris=120, seed=31)
    oc, od, _ = scenes.run_oracle(scene)
    lib = capi.load_oracle()
    # oracle: draw the same geometry a second time on top (LESS_OR_EQUAL: equal depths pass again)
    mem = scenes.HostMemory(); m = scenes.materialize(scene, mem.alloc)
    for img, att in ((scene.color, m.color_attachment), (scene.depth, m.depth_attachment)):
        cv, is_ds = scenes.clear_value(img)
        import ctypes as C
        assert lib.cpvk_oracle_clear(C.byref(att), C.byref(cv), is_ds) == 0
    for _ in range(2):
        assert lib.cpvk_oracle_draw(C.byref(m.desc), C.byref(m.state), None) == 0
    s = SceneOnDevice(group, scene)
    try:
        s.clear(); s.draw(); s.draw(); s.gather()
        assert np.array_equal(s.read_color(), mem.arrays["color"][:scene.color.nbytes])
        assert np.array_equal(s.read_depth(), mem.arrays["depth"][:scene.depth.nbytes])
    finally:
        s.close()


def test_group_transfer_commands_are_replicated(group):
    """Clears, copies and blits run on every member: afterwards any member's replica can feed a draw, and the leader's is read."""
    import ctypes as C
    lib = capi.load_oracle()
    rng = np.random.default_rng(5)
    w, h = 64, 40
    src_host = rng.integers(0, 256, w * h * 4, dtype=np.uint8)
    dst_host = np.zeros(2 * w * 2 * h * 8, dtype=np.uint8)
    src = group.alloc(src_host.nbytes); dst = group.alloc(dst_host.nbytes)
    group.upload(src, src_host); group.upload(dst, dst_host)
    hb = capi.Blit(capi.Attachment(src_host.ctypes.data, w, h, w * 4, 37), capi.Attachment(dst_host.ctypes.data, 2 * w, 2 * h, 2 * w * 8, 97), 0, 0, w, h, 0, 0, 2 * w, 2 * h, 1)
    db = capi.Blit(capi.Attachment(src, w, h, w * 4, 37), capi.Attachment(dst, 2 * w, 2 * h, 2 * w * 8, 97), 0, 0, w, h, 0, 0, 2 * w, 2 * h, 1)
    assert lib.cpvk_oracle_blit(C.byref(hb)) == 0
    group.blit(db)
    assert np.array_equal(group.download(dst, dst_host.nbytes), dst_host)
    group.free(src); group.free(dst)


@pytest.mark.parametrize("devices", ["0,0", "all"])
def test_icd_on_a_group(tmp_path, built, devices):
    """The Vulkan ICD with CPVK_CUDA_DEVICES: vkCreateDevice builds the group, every vkCmdDraw* is split into bands,
    vkCmdEndRenderPass exchanges them, vkQueueSubmit returns with the whole frame readable."""
    if devices == "all":
        n = gpu_count()
        if n < 2:
            pytest.skip("one GPU visible")
        devices = ",".join(str(i) for i in range(min(n, 8)))
    for scene in (scenes.draw_cube(), scenes.draw_textured_cube(filt=scenes.LINEAR), scenes.mesh_indexed(640, 360, 160, 90)):
        oc, od, _ = scenes.run_oracle(scene)
        gc, gd, info = scenes.run_icd(scene, str(tmp_path / scene.name), frames=2, env={"CPVK_CUDA_DEVICES": devices})
        assert np.array_equal(gc, oc), scene.name
        if od is not None and gd is not None:
            assert np.array_equal(gd, od), scene.name


def run_barrier_pair(tmp_path, scenario):
    """Two PROCESSES on GPU 0 (tests/peer_barrier_worker.py), the way one-process-per-GPU runs use the barrier: flag arrays
    and data reach the other side as cudaIpc mappings. (Two streams of one process would do on paper, but a kernel that
    spins on a flag may sit in front of the other stream's work in a shared hardware queue; separate processes have their
    own queues and the GPU time-slices between them.)"""
    import os
    import subprocess
    import sys
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "peer_barrier_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, role, scenario, str(tmp_path)], stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
             for role in ("producer", "consumer")]
    outs = []
    try:
        for pr in procs:
            out, _ = pr.communicate(timeout=240)
            outs.append(out)
    finally:
        for pr in procs:
            if pr.poll() is None:
                pr.kill()
    for pr, out in zip(procs, outs):
        assert pr.returncode == 0, out[-3000:]
    assert "consumer ok" in outs[1], outs[1][-3000:]


def test_peer_barrier_orders_two_processes(built, tmp_path):
    """cpvk_cuda_peer_barrier (the per-frame ordering step of one-process-per-GPU runs): a participant's work behind the barrier
    runs after every participant's work in front of it — the consumer copies what the producer wrote, eight rounds, each with
    new bytes."""
    run_barrier_pair(tmp_path, "copy")


def test_peer_barrier_behind_a_draw_that_is_replayed(built, tmp_path):
    """A barrier called right behind a draw does not wait for the draw's validation: it reads the verdict on the device. The
    producer's first draw does not fit the launch plan of a fresh device (900 triangles over four tiles: more than the 512
    list slots of single-pass binning), so its tail and the barriers behind it are no-ops, and all are issued again when the
    next call looks at the verdict. The consumer copies the producer's frame behind the barrier: it must be the finished
    one (the oracle's bytes), in the replayed round and in the rounds that run the learned plan."""
    run_barrier_pair(tmp_path, "draw")

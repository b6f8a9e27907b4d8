"""Tiled multi-GPU frame, exchanged by an NCCL all-gather and by the peer-memory stores fused into
the gather kernel: needs >= 2 CUDA devices."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, W, H, q):
    import torch.distributed as dist
    import radiancecascade_b200 as rc
    from radiancecascade_b200 import _ffi, distributed as rd
    from common import frame_setup
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    st, _, _ = frame_setup("teapot", W, H)
    tr = rd.TiledRenderer(rank, world, rank, (W, H), st, rc.scenes.scene_path("teapot"))
    tr.render(st)
    full = tr.gather()
    ok = True
    want = None
    if rank == 0:
        ref = rc.DefaultRenderer.new(0, (W, H), st, rc.scenes.scene_path("teapot"))
        ref.update(st); ref.render()
        want = ref.read_target(_ffi.RC_TARGET_IRRADIANCE)
        ok = bool(np.array_equal(full.cpu().numpy().view(np.uint16), want.view(np.uint16)))
    # the same frame through the peer-memory exchange fused into the gather kernel (no collective): three frames,
    # so that both buffer slots and the release / arrive handshake are exercised
    tr.attach_peers()
    for _ in range(3):
        tr.render(st)
        peer_full = tr.gather_peer().clone()
    torch.cuda.synchronize()
    _, _, timeouts = tr.renderer.peer_frame(check=True)
    if rank == 0:
        ok = ok and timeouts == 0 and bool(np.array_equal(peer_full.cpu().numpy().view(np.uint16), want.view(np.uint16)))
        q.put(ok)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_tiled_frame_equals_single_gpu():
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 384, 216, q)) for r in range(2)]
    for p in procs:
        p.start()
    ok = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
    assert ok

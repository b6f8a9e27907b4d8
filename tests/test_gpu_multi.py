"""Tiled multi-GPU frame, exchanged by an NCCL all-gather and by the peer-memory stores fused into the gather
kernel, with the strip cuts moving between frames (rc_set_tile): needs >= 2 CUDA devices.  Every frame uses a different
camera and is compared with its own single-GPU frame, so a stale buffer slot, an early read or a broken arrive / release
handshake cannot hide behind identical pixels."""
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
    name = "teapot"
    st, _, _ = frame_setup(name, W, H, frame=0)
    tr = rd.TiledRenderer(rank, world, rank, (W, H), st, rc.scenes.scene_path(name), balance=True)
    ref = rc.DefaultRenderer.new(rank, (W, H), st, rc.scenes.scene_path(name)) if rank == 0 else None
    stream = torch.cuda.Stream()
    ok, why = True, ""

    def want_frame(state):
        ref.update(state); ref.render()
        return ref.read_target(_ffi.RC_TARGET_IRRADIANCE).view(np.uint16)

    # (1) NCCL all-gather of the strips
    tr.render(st)
    full = tr.gather()
    if rank == 0 and not np.array_equal(full.cpu().numpy().view(np.uint16), want_frame(st)):
        ok, why = False, "nccl all-gather frame differs"
    # (2) the peer-memory exchange fused into the gather kernel: six frames, each with its own camera, the cuts moved twice
    tr.attach_peers()
    for f in range(6):
        if f == 4:      # frames 0-3: final image gather on rank 0 (default); frames 4-5: every rank receives the frame
            tr.renderer.set_tuning("peer_broadcast", 1)
        stf, _, _ = frame_setup(name, W, H, frame=3 + 5 * f)
        with torch.cuda.stream(stream):
            tr.render(stf, stream.cuda_stream)
            got = tr.gather_peer(stream.cuda_stream).clone()       # consumed on the render stream, before the next render
        stream.synchronize()
        if rank == 0 and not np.array_equal(got.cpu().numpy().view(np.uint16), want_frame(stf)):
            ok, why = False, f"peer frame {f} differs"
        if f >= 4:      # all-gather mode: every rank's copy equals rank 0's
            mine = got.contiguous()
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            if rank == 0 and not all(torch.equal(p, parts[0]) for p in parts):
                ok, why = False, f"broadcast frame {f}: ranks disagree"
        if f in (1, 3):      # uneven "measured" times: every rank computes the same new cuts and re-tiles in place
            times = [1.0 + 0.8 * r for r in range(world)] if f == 1 else [1.6 - 0.5 * r for r in range(world)]
            moved = tr.rebalance(tr.all_gather_times(times[rank]))
            if not moved:
                ok, why = False, "the balancer did not move the cuts"
    torch.cuda.synchronize()
    # (3) halo EXCHANGE (RC_CFG_HALO_EXCHANGE): levels >= 1 marched only by the owners, request masks and child averages moved
    # with NCCL send / recv — the assembled frame is still the single-GPU frame bit for bit, fewer rays are marched than with
    # recomputed halos, and it survives moving the cuts
    trx = rd.TiledRenderer(rank, world, rank, (W, H), st, rc.scenes.scene_path(name), balance=True, halo_exchange=True)
    marched_x = marched_r = 0
    for f in range(4):
        stf, _, _ = frame_setup(name, W, H, frame=2 + 7 * f)
        with torch.cuda.stream(stream):
            trx.render(stf, stream.cuda_stream)
            tr.render(stf, stream.cuda_stream)
        stream.synchronize()
        marched_x += sum(m for m in trx.renderer.rays_marched() if m)
        marched_r += sum(m for m in tr.renderer.rays_marched() if m)
        fullx = rd.all_gather_tiles(rd.irradiance_tensor(trx.renderer), trx.tiles, W, H)
        if rank == 0 and not np.array_equal(fullx.cpu().numpy().view(np.uint16), want_frame(stf)):
            ok, why = False, f"halo-exchange frame {f} differs"
        if f == 1:
            times = [1.0 + 0.7 * r for r in range(world)]
            if not trx.rebalance(trx.all_gather_times(times[rank])):
                ok, why = False, "the balancer did not move the cuts (exchange)"
    tot = torch.tensor([float(marched_x), float(marched_r)], device="cuda", dtype=torch.float64)
    dist.all_reduce(tot)
    if rank == 0 and not (tot[0] < tot[1]):
        ok, why = False, f"exchange marched {tot[0].item()} rays, recompute {tot[1].item()}"
    torch.cuda.synchronize()
    _, _, timeouts = tr.renderer.peer_frame(check=True)
    if rank == 0:
        if timeouts != 0:
            ok, why = False, f"{timeouts} flag time-outs"
        q.put((ok, why))
    dist.barrier()
    dist.destroy_process_group()


def _run(world, W, H):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, W, H, q)) for r in range(world)]
    for p in procs:
        p.start()
    ok, why = q.get(timeout=600)
    for p in procs:
        p.join(timeout=120)
    assert ok, why


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_gpu_tiled_frame_equals_single_gpu():
    _run(2, 384, 216)


@pytest.mark.skipif(torch.cuda.device_count() < 4, reason="needs 4 GPUs")
def test_four_gpu_tiled_frame_equals_single_gpu():
    _run(4, 640, 360)


@pytest.mark.skipif(torch.cuda.device_count() < 8, reason="needs 8 GPUs")
def test_eight_gpu_tiled_frame_equals_single_gpu():
    _run(8, 960, 544)

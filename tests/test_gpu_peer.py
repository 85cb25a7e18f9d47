"""-m gpu: the peer-memory collectives of the sharded step (csrc/peer_collectives.cuh) exercised on ONE
device: `world` ranks are emulated by `world` CUDA streams, each with its own data block, flag area
and epoch array (all ordinary device memory here; symmetric peer-mapped memory on a multi-GPU box).
Sizes are chosen so that the grids of all emulated ranks are co-resident on one device (<= ~250 CTAs).
The kernels of the emulated ranks run concurrently and synchronise through the same flag protocol
(st.release.sys / ld.acquire.sys), so barrier logic, epochs across repeated launches, slice
partitioning and segment addressing are covered without a second GPU.  The multi-GPU runs are in
tests/test_gpu_sharded.py."""
import ctypes

import pytest
import torch

pytestmark = pytest.mark.gpu

TIMEOUT_MS = 4000


def _setup(world, nbytes):
    from multimodal_baby_b200 import _cabi
    lib = _cabi.load()
    dev = torch.device("cuda:0")
    fw, nblk = int(lib.cvcl_peer_flag_words()), int(lib.cvcl_peer_max_blocks())
    data = [torch.empty(nbytes, dtype=torch.uint8, device=dev) for _ in range(world)]
    flags = [torch.zeros(fw, dtype=torch.int32, device=dev) for _ in range(world)]
    epoch = [torch.zeros(nblk, dtype=torch.int32, device=dev) for _ in range(world)]
    status = [torch.zeros(1, dtype=torch.int32, device=dev) for _ in range(world)]
    arr = ctypes.c_void_p * world
    streams = [torch.cuda.Stream(device=dev) for _ in range(world)]
    return _cabi, dev, data, flags, epoch, status, arr, streams


def _join(streams, status):
    for s in streams:
        s.synchronize()
    assert all(int(t.item()) == 0 for t in status), "a cross-rank barrier timed out"


@pytest.mark.parametrize("world,n", [(2, 8 + 1024), (2, 2_400_012), (4, 400_012), (8, 400_012), (8, 8)])
def test_peer_allreduce_emulated_ranks(world, n):
    """in-place two-shot sum == sum of the rank buffers in rank order, bit-identical on every rank,
    three launches in a row (epochs advance, flags are never reset)."""
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, n * 4)
    p_data = arr(*[d.data_ptr() for d in data]); p_flags = arr(*[f.data_ptr() for f in flags])
    g = torch.Generator(device="cpu").manual_seed(world * 1000 + n % 997)
    for it in range(3):
        src = [torch.randn(n, generator=g) for _ in range(world)]
        for d, s in zip(data, src):
            d.view(torch.float32).copy_(s.to(dev))
        torch.cuda.synchronize()
        for r in range(world):
            _cabi.call("cvcl_peer_allreduce_f32", p_data, p_flags, epoch[r].data_ptr(), status[r].data_ptr(),
                       world, r, n, TIMEOUT_MS, streams[r].cuda_stream)
        _join(streams, status)
        want = src[0].clone()
        for s in src[1:]:
            want += s                      # fp32, rank order
        for r in range(world):
            assert torch.equal(data[r].view(torch.float32).cpu(), want), (world, n, it, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_allgather_emulated_ranks(world):
    """feature exchange (1 segment) and LSE exchange (2 segments -> [2, world*b]) against torch.cat."""
    b, E = 64, 512
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, b * 2 * E * 2)
    p_data = arr(*[d.data_ptr() for d in data]); p_flags = arr(*[f.data_ptr() for f in flags])
    g = torch.Generator(device="cpu").manual_seed(7 + world)
    for it in range(2):
        blocks = [torch.randn(b, 2 * E, generator=g).to(torch.bfloat16) for _ in range(world)]
        for d, s in zip(data, blocks):
            d.view(torch.bfloat16).view(b, 2 * E).copy_(s.to(dev))
        dst = [torch.zeros(world * b, 2 * E, dtype=torch.bfloat16, device=dev) for _ in range(world)]
        torch.cuda.synchronize()
        nbytes = b * 2 * E * 2
        for r in range(world):
            _cabi.call("cvcl_peer_allgather", p_data, p_flags, epoch[r].data_ptr(), status[r].data_ptr(), world, r,
                       nbytes, 1, 0, dst[r].data_ptr(), 0, TIMEOUT_MS, streams[r].cuda_stream)
        _join(streams, status)
        want = torch.cat(blocks)
        for r in range(world):
            assert torch.equal(dst[r].cpu(), want), (world, it, r)
    # two segments per rank (lse0 | lse1) -> [2, world*b]
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, 2 * b * 4)
    p_data = arr(*[d.data_ptr() for d in data]); p_flags = arr(*[f.data_ptr() for f in flags])
    lses = [torch.randn(2, b, generator=g) for _ in range(world)]
    for d, s in zip(data, lses):
        d.view(torch.float32).view(2, b).copy_(s.to(dev))
    dst = [torch.zeros(2, world * b, device=dev) for _ in range(world)]
    torch.cuda.synchronize()
    for r in range(world):
        _cabi.call("cvcl_peer_allgather", p_data, p_flags, epoch[r].data_ptr(), status[r].data_ptr(), world, r,
                   b * 4, 2, b * 4, dst[r].data_ptr(), world * b * 4, TIMEOUT_MS, streams[r].cuda_stream)
    _join(streams, status)
    want = torch.cat(lses, dim=1)
    for r in range(world):
        assert torch.equal(dst[r].cpu(), want), (world, r)


def test_peer_barrier_orders_streams():
    """rank 1 writes after a delay, both pass the barrier, rank 0 then reads rank 1's value."""
    world = 2
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, 16)
    p_flags = arr(*[f.data_ptr() for f in flags])
    out = torch.zeros(1, dtype=torch.float32, device=dev)
    val = data[1].view(torch.float32)
    val.zero_()
    torch.cuda.synchronize()
    with torch.cuda.stream(streams[1]):
        torch.cuda._sleep(20_000_000)                # ~10 ms at 2 GHz
        val.fill_(42.0)
    for r in range(world):
        _cabi.call("cvcl_peer_barrier", p_flags, epoch[r].data_ptr(), status[r].data_ptr(), world, r, TIMEOUT_MS,
                   streams[r].cuda_stream)
    with torch.cuda.stream(streams[0]):
        out.copy_(val[:1])
    _join(streams, status)
    assert float(out.item()) == 42.0


@pytest.mark.parametrize("world,n", [(2, 8 + 1024), (2, 2_400_012), (4, 400_012), (8, 400_012), (8, 8)])
def test_peer_allreduce_push_emulated_ranks(world, n):
    """push variant (scatter into the owners' scratch blocks, local reduce, broadcast): same contract."""
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, n * 4)
    nscr = int(_cabi.load().cvcl_peer_allreduce_scratch_bytes(n, world))
    scratch = [torch.empty(nscr, dtype=torch.uint8, device=dev) for _ in range(world)]
    p_data = arr(*[d.data_ptr() for d in data]); p_flags = arr(*[f.data_ptr() for f in flags])
    p_scr = arr(*[d.data_ptr() for d in scratch])
    g = torch.Generator(device="cpu").manual_seed(world * 1000 + n % 997 + 1)
    for it in range(3):
        src = [torch.randn(n, generator=g) for _ in range(world)]
        for d, s in zip(data, src):
            d.view(torch.float32).copy_(s.to(dev))
        torch.cuda.synchronize()
        for r in range(world):
            _cabi.call("cvcl_peer_allreduce_push_f32", p_data, p_scr, p_flags, epoch[r].data_ptr(),
                       status[r].data_ptr(), world, r, n, TIMEOUT_MS, streams[r].cuda_stream)
        _join(streams, status)
        want = src[0].clone()
        for s in src[1:]:
            want += s
        for r in range(world):
            assert torch.equal(data[r].view(torch.float32).cpu(), want), (world, n, it, r)


@pytest.mark.parametrize("world", [2, 4, 8])
def test_peer_allgather_push_emulated_ranks(world):
    b, E = 64, 512
    _cabi, dev, _, flags, epoch, status, arr, streams = _setup(world, 16)
    p_flags = arr(*[f.data_ptr() for f in flags])
    g = torch.Generator(device="cpu").manual_seed(70 + world)
    dst = [torch.zeros(world * b, 2 * E, dtype=torch.bfloat16, device=dev) for _ in range(world)]
    p_dst = arr(*[d.data_ptr() for d in dst])
    nbytes = b * 2 * E * 2
    for it in range(2):
        blocks = [torch.randn(b, 2 * E, generator=g).to(torch.bfloat16).to(dev) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            _cabi.call("cvcl_peer_allgather_push", p_dst, p_flags, epoch[r].data_ptr(), status[r].data_ptr(), world, r,
                       blocks[r].data_ptr(), nbytes, 1, 0, 0, TIMEOUT_MS, streams[r].cuda_stream)
        _join(streams, status)
        want = torch.cat(blocks).cpu()
        for r in range(world):
            assert torch.equal(dst[r].cpu(), want), (world, it, r)
    # two segments per rank (lse0 | lse1) -> [2, world*b]
    _cabi, dev, _, flags, epoch, status, arr, streams = _setup(world, 16)
    p_flags = arr(*[f.data_ptr() for f in flags])
    lses = [torch.randn(2, b, generator=g).to(dev) for _ in range(world)]
    dst = [torch.zeros(2, world * b, device=dev) for _ in range(world)]
    p_dst = arr(*[d.data_ptr() for d in dst])
    torch.cuda.synchronize()
    for r in range(world):
        _cabi.call("cvcl_peer_allgather_push", p_dst, p_flags, epoch[r].data_ptr(), status[r].data_ptr(), world, r,
                   lses[r].data_ptr(), b * 4, 2, b * 4, world * b * 4, TIMEOUT_MS, streams[r].cuda_stream)
    _join(streams, status)
    want = torch.cat(lses, dim=1).cpu()
    for r in range(world):
        assert torch.equal(dst[r].cpu(), want), (world, r)


@pytest.mark.parametrize("push", [False, True])
def test_peer_allreduce_alternating_sizes_share_a_channel(push):
    """A training step (large all-reduce, many CTAs) followed by forward-only steps (n = 8, one CTA)
    and a training step again on the SAME channel: flag words are indexed with a fixed per-phase stride, so
    the single-CTA launches cannot leave a high epoch in a word that the many-CTA grid reads as another
    block's (slot = (phase * gridDim.x + block) aliased phase 1 / block 0 onto phase 0 /
    block 1)."""
    world, big, small = 2, 1_200_000, 8
    _cabi, dev, data, flags, epoch, status, arr, streams = _setup(world, big * 4)
    nscr = int(_cabi.load().cvcl_peer_allreduce_scratch_bytes(big, world))
    scratch = [torch.empty(nscr, dtype=torch.uint8, device=dev) for _ in range(world)]
    p_data = arr(*[d.data_ptr() for d in data]); p_flags = arr(*[f.data_ptr() for f in flags])
    p_scr = arr(*[d.data_ptr() for d in scratch])
    g = torch.Generator(device="cpu").manual_seed(4242)
    for it, n in enumerate([big, small, small, small, big, small, big]):
        src = [torch.randn(n, generator=g) for _ in range(world)]
        for d, s in zip(data, src):
            d.view(torch.float32)[:n].copy_(s.to(dev))
        torch.cuda.synchronize()
        for r in range(world):
            if push:
                _cabi.call("cvcl_peer_allreduce_push_f32", p_data, p_scr, p_flags, epoch[r].data_ptr(),
                           status[r].data_ptr(), world, r, n, TIMEOUT_MS, streams[r].cuda_stream)
            else:
                _cabi.call("cvcl_peer_allreduce_f32", p_data, p_flags, epoch[r].data_ptr(), status[r].data_ptr(),
                           world, r, n, TIMEOUT_MS, streams[r].cuda_stream)
        _join(streams, status)
        want = src[0] + src[1]
        for r in range(world):
            assert torch.equal(data[r].view(torch.float32)[:n].cpu(), want), (it, n, r)

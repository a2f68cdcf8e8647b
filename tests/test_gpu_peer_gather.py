"""The NVLink peer-memory all-gather (pipeline.PeerGather) against NCCL's all_gather_into_tensor, world size 2..N on one
node.  Needs at least two GPUs: skipped on single-GPU boxes (run with `gpurun --gpus 2 -- python -m pytest
tests/test_gpu_peer_gather.py -m gpu`)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_words, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", device_id=dev)
    try:
        from densematcher_b200 import pipeline
        pg = pipeline.PeerGather(n_words, dev)
        sends = [torch.empty(n_words, dtype=torch.int32, device=dev) for _ in range(2)]
        ok = True
        for it in range(7):                                        # both buffer sets, several reuses
            s = sends[it & 1]
            torch.cuda.current_stream().wait_event(pg.done[(pg.i ^ 1)])   # the gather that read this send buffer two steps ago
            g = torch.Generator(device=dev).manual_seed(1000 * it + rank)
            s.copy_(torch.randint(-2**31, 2**31 - 1, (n_words,), dtype=torch.int64, device=dev, generator=g).to(torch.int32))
            ref = torch.empty(world * n_words, dtype=torch.int32, device=dev)
            dist.all_gather_into_tensor(ref, s)
            got = pg.gather(s)
            pg.wait()
            ok = ok and bool(torch.equal(got, ref))
        q.put((rank, ok))
    finally:
        dist.destroy_process_group()


def test_peer_gather_matches_nccl():
    world = min(torch.cuda.device_count(), 8)
    if world < 2:
        pytest.skip("needs at least two GPUs on the node")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, 3_500_001, q)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
    res = sorted(q.get(timeout=5) for _ in range(world))
    assert res == [(r, True) for r in range(world)], res

"""CPU, world_size 2, gloo: the N > 1 plumbing — one broadcast of the model image at init, round-robin
sharding, no collective in steady state, order restored by the launcher.  The per-rank analyzer here is the
oracle (the CUDA path needs a GPU); the sharding/merge code under test is the product's."""
import os, socket, sys
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, image_path, texts, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from kiwi_b200 import shard
    from tests.orc import Oracle
    image = open(image_path, "rb").read(4096) if rank == 0 else b""      # header page is enough to check the broadcast
    got = shard.broadcast_image(image, dist)
    ok_image = len(got) == 4096 and got[:7] == b"KB2IMG1"
    o = Oracle(image_path)
    mine = shard.shard_indices(len(texts), rank, world)
    res = [o.analyze(texts[i]) for i in mine]
    q.put((rank, ok_image, res))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    from tests.orc import IMAGE, Oracle
    if not os.path.exists(IMAGE):
        pytest.skip("model image missing")
    from tests.goldenio import read_inputs
    from kiwi_b200 import shard
    texts = read_inputs("inputs_web")[:24]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, IMAGE, texts, q)) for r in range(2)]
    for p in procs: p.start()
    got = {}
    for _ in range(2):
        rank, ok_image, res = q.get(timeout=240)
        assert ok_image
        got[rank] = res
    for p in procs: p.join(timeout=60)
    merged = shard.merge_round_robin([got[0], got[1]], len(texts))
    o = Oracle(IMAGE)
    assert merged == [o.analyze(t) for t in texts]

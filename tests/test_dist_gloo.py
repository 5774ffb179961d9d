"""N > 1 path on CPU: two gloo ranks shard a global batch the way bench.py / the multi-GPU predict path do
(contiguous split, no data-path collective), time a step, take the max over ranks and gather the per-image
detection counts back in global order."""
import os
import sys
from pathlib import Path

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1",
                      MASTER_PORT=str(port))
    sys.path.insert(0, str(ROOT / "spectrogram-yolov11_b200"))
    from specyolo import dist as sdist
    from oracle import nms_ref

    r, w = sdist.init("gloo")
    assert (r, w) == (rank, world)
    n_global = 7                                            # ragged: 4 + 3
    lo, hi = sdist.shard_range(n_global, r, w)
    g = torch.Generator().manual_seed(5)
    pred = torch.rand((n_global, 6, 64), generator=g)      # same global batch on every rank, each takes its slice
    pred[:, 2:4] = pred[:, 2:4] * 50 + 5
    pred[:, :2] *= 200
    mine = pred[lo:hi]
    sdist.barrier()
    out = nms_ref.non_max_suppression(mine.numpy(), 0.5, 0.5)     # stand-in for the per-rank replica (CPU test)
    counts = torch.tensor([len(o) for o in out], dtype=torch.int64)
    t = sdist.max_over_ranks(0.1 * (rank + 1))
    total = sdist.sum_over_ranks(float(hi - lo))
    allc = sdist.gather_counts(counts)
    # the product's sharded predict API: one global batch in, all detections back in input order on every rank
    fn = lambda shard: [torch.from_numpy(o) for o in nms_ref.non_max_suppression(shard.numpy(), 0.5, 0.5)]
    full = sdist.predict_sharded(fn, pred)
    tiny = sdist.predict_sharded(fn, pred[:1])              # fewer units than ranks: rank 1's shard is empty
    sdist.barrier()
    ref = nms_ref.non_max_suppression(pred.numpy(), 0.5, 0.5)
    assert len(full) == n_global and all(torch.equal(a, torch.from_numpy(b)) for a, b in zip(full, ref))
    assert len(tiny) == 1 and torch.equal(tiny[0], torch.from_numpy(ref[0]))
    if rank == 0:
        q.put((t, total, torch.cat(allc).tolist(), [len(o) for o in ref]))
    dist.destroy_process_group()


def test_two_rank_sharding_gloo():
    sys.path.insert(0, str(ROOT))
    ctx = mp.get_context("spawn")
    q = ctx.SimpleQueue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t, total, gathered, ref = q.get()
    assert abs(t - 0.2) < 1e-9            # max over ranks
    assert total == 7.0                   # every image processed exactly once
    assert gathered == ref                # sharded result == unsharded result, in global order

"""Frame sharding across the GPUs of one box.  Frames are independent (the reference's paf_processor
keeps no state between frames), so ranks take contiguous frame ranges, run their own handle on their
own GPU, and the skeletons are gathered on the host in frame order.  There is no collective on the
data path; torch.distributed (nccl on GPUs, gloo in CPU tests) only carries the final host gather."""
import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous [start, stop) of rank; sizes differ by at most one frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def gather_results(local, rank, world, group=None):
    """local = (humans [n_local, max_humans] records, counts [n_local], flags [n_local]).
    Returns the concatenation over ranks in frame order on rank 0, None elsewhere."""
    if world == 1:
        return local
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(local, bucket, dst=0, group=group)
    if rank != 0:
        return None
    return tuple(np.concatenate([b[i] for b in bucket]) for i in range(3))


def process_stream(engine, conf, paf, rank=0, world=1, batch=None, group=None, **kw):
    """Runs this rank's shard of a stream of frames through `engine` with all its slots in flight and
    gathers on rank 0.  conf/paf hold the WHOLE stream (numpy or CUDA tensors); only the shard is read."""
    from . import _capi as capi
    n = int(conf.shape[0])
    lo, hi = shard_range(n, rank, world)
    batch = batch or engine.max_batch
    humans = np.zeros((hi - lo, engine.max_humans), capi.HUMAN_DT)
    counts = np.zeros(hi - lo, np.int32)
    flags = np.zeros(hi - lo, np.int32)
    inflight = []
    n_slots = int(engine.cfg.n_slots)
    for s in range(lo, hi, batch):
        e = min(s + batch, hi)
        if len(inflight) == n_slots:
            engine.wait(inflight.pop(0))
        out = (humans[s - lo:e - lo], counts[s - lo:e - lo], flags[s - lo:e - lo])
        inflight.append(engine.submit(conf[s:e], paf[s:e], out=out, **kw))
    for t in inflight:
        engine.wait(t)
    return gather_results((humans, counts, flags), rank, world, group)


def _run_shard(engine, conf, paf, lo, hi, batch, outs, kw):
    """Keeps all slots of one engine busy over frames [lo, hi); results land in the caller's arrays."""
    humans, counts, flags = outs
    inflight = []
    n_slots = int(engine.cfg.n_slots)
    for s in range(lo, hi, batch):
        e = min(s + batch, hi)
        if len(inflight) == n_slots:
            engine.wait(inflight.pop(0))
        inflight.append(engine.submit(conf[s:e], paf[s:e], out=(humans[s:e], counts[s:e], flags[s:e]), **kw))
    for t in inflight:
        engine.wait(t)


def process_stream_multi(engines, conf, paf, batch=None, **kw):
    """One process, several GPUs (BASELINE.json configs[4]: a stream sharded over the GPUs of one box,
    per-GPU streams, host gather).  `engines` = one Engine per device; the stream is cut into contiguous
    shards, one host thread per GPU keeps all slots of its engine in flight (the C-ABI calls release the
    GIL and a blocking wait on one GPU never stalls submission to another), and the results come back in
    frame order.  conf/paf are host arrays (pinned for full PCIe rate)."""
    import threading
    from . import _capi as capi
    n, world = int(conf.shape[0]), len(engines)
    outs = (np.zeros((n, engines[0].max_humans), capi.HUMAN_DT), np.zeros(n, np.int32), np.zeros(n, np.int32))
    errors, threads = [], []

    def worker(r):
        try:
            lo, hi = shard_range(n, r, world)
            _run_shard(engines[r], conf, paf, lo, hi, batch or engines[r].max_batch, outs, kw)
        except BaseException as e:  # re-raised on the calling thread
            errors.append(e)

    for r in range(1, world):
        threads.append(threading.Thread(target=worker, args=(r,)))
        threads[-1].start()
    worker(0)
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return outs

"""Frame sharding across the GPUs of one box.  Frames are independent (the reference's paf_processor
keeps no state between frames), so ranks take contiguous frame ranges, run their own handle on their
own GPU, and the skeletons are gathered on the host in frame order.  There is no collective on the
data path.  Two host gathers:

* `HostGather` - the ranks of ONE box map one POSIX shared-memory segment that holds the result arrays of the whole
  stream; every rank pins its own slice (opp_host_register) so that its GPU's assembly kernel writes the skeletons
  straight into the shared buffer.  "Gathering" is then a sequence-number handshake in the same segment: no copy,
  no serialisation, a few microseconds.  This is what bench.py times for BASELINE.json configs[4].
* `gather_results` - torch.distributed gather_object of compacted records (gloo in CPU tests): works across boxes.
"""
import os
import time

import numpy as np


def shard_range(n_frames, rank, world):
    """Contiguous [start, stop) of rank; sizes differ by at most one frame."""
    base, rem = divmod(n_frames, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def _compact(local):
    """(humans [n, cap], counts, flags) -> (records of the humans actually found, counts, flags, cap): what travels."""
    humans, counts, flags = local
    cap = humans.shape[1] if humans.ndim == 2 else 0
    keep = np.arange(cap)[None, :] < np.minimum(counts, cap)[:, None] if len(counts) else np.zeros((0, cap), bool)
    return humans[keep], counts, flags, cap


def _expand(packed):
    recs, counts, flags, cap = packed
    humans = np.zeros((len(counts), cap), recs.dtype)
    keep = np.arange(cap)[None, :] < np.minimum(counts, cap)[:, None] if len(counts) else np.zeros((0, cap), bool)
    humans[keep] = recs
    return humans, counts, flags


def gather_results(local, rank, world, group=None):
    """local = (humans [n_local, max_humans] records, counts [n_local], flags [n_local]).
    Returns the concatenation over ranks in frame order on rank 0, None elsewhere.  Only the humans actually
    found travel (a 128-slot frame is 37 KB, its five humans 1.5 KB)."""
    if world == 1:
        return local
    import torch.distributed as dist
    bucket = [None] * world if rank == 0 else None
    dist.gather_object(_compact(local), bucket, dst=0, group=group)
    if rank != 0:
        return None
    parts = [_expand(b) for b in bucket]
    return tuple(np.concatenate([p[i] for p in parts]) for i in range(3))


class HostGather:
    """Result arrays of a whole stream in one shared-memory segment mapped by every rank of the box.

        g = HostGather(name, n_frames, max_humans, rank, world)      # rank 0 creates, the others attach
        humans, counts, flags = g.local()                            # this rank's slice: pass as Engine `out=` buffers
        ... process the shard ...
        g.publish()                                                  # this rank's results are in place
        whole = g.collect()                                          # rank 0: waits for every rank, returns the full arrays

    Layout: [world] int64 sequence numbers (one cache line each) | humans [n, max_humans] x 292 B | counts [n] | flags [n],
    every array starting on a page boundary.  `register=True` pins this rank's slices with opp_host_register so the
    GPU writes them directly; without a GPU (CPU tests) the arrays are plain shared memory."""
    PAGE = 4096
    LINE = 64

    def __init__(self, name, n_frames, max_humans, rank, world, register=True, timeout=120.0):
        import mmap
        from . import _capi as capi
        self.rank, self.world, self.n, self.cap, self.timeout = rank, world, n_frames, max_humans, timeout
        up = lambda v: (v + self.PAGE - 1) // self.PAGE * self.PAGE
        self.off_h = up(2 * world * self.LINE)
        self.off_c = self.off_h + up(n_frames * max_humans * capi.HUMAN_DT.itemsize)
        self.off_f = self.off_c + up(n_frames * 4)
        size = self.off_f + up(n_frames * 4)
        # POSIX shared memory by hand (/dev/shm + mmap): the segment must outlive numpy views handed to callers and
        # must not be unlinked by the resource tracker of whichever rank exits first
        self.path = os.path.join("/dev/shm", name)
        if rank == 0:
            try:  # a crashed earlier run may have left the name behind
                os.unlink(self.path)
            except FileNotFoundError:
                pass
            fd = os.open(self.path + ".tmp", os.O_CREAT | os.O_RDWR | os.O_TRUNC, 0o600)
            os.ftruncate(fd, size)                      # zero pages: all sequence numbers start at 0
            os.rename(self.path + ".tmp", self.path)    # visible to the other ranks only at its full size
        else:
            t0 = time.monotonic()
            while True:
                try:
                    fd = os.open(self.path, os.O_RDWR)
                    break
                except FileNotFoundError:
                    if time.monotonic() - t0 > timeout:
                        raise TimeoutError("HostGather: rank 0 never created %r" % self.path)
                    time.sleep(0.002)
        try:
            buf = mmap.mmap(fd, size)
        finally:
            os.close(fd)
        self._map = buf
        self._seq = np.frombuffer(buf, np.int64, 2 * world * self.LINE // 8)
        self.humans = np.frombuffer(buf, capi.HUMAN_DT, n_frames * max_humans, self.off_h).reshape(n_frames, max_humans)
        self.counts = np.frombuffer(buf, np.int32, n_frames, self.off_c)
        self.flags = np.frombuffer(buf, np.int32, n_frames, self.off_f)
        self.lo, self.hi = shard_range(n_frames, rank, world)
        self._round = 0
        self._registered = []
        if register and self.hi > self.lo:
            import ctypes as C
            L = capi.lib()
            base = C.addressof(C.c_char.from_buffer(buf))
            rec = max_humans * capi.HUMAN_DT.itemsize
            for a, b in ((self.off_h + self.lo * rec, self.off_h + self.hi * rec), (self.off_c + self.lo * 4, self.off_c + self.hi * 4),
                         (self.off_f + self.lo * 4, self.off_f + self.hi * 4)):
                a, b = a // self.PAGE * self.PAGE, up(b)   # whole pages (a boundary page may also be pinned by the neighbour rank)
                if L.opp_host_register(base + a, b - a) != capi.OK:
                    raise capi.OppError(capi.ERR_CUDA, (L.opp_last_error(None) or b"").decode())
                self._registered.append(base + a)

    def local(self):
        """This rank's slices of (humans, counts, flags): hand them to Engine.submit(out=...) / process_stream."""
        return self.humans[self.lo:self.hi], self.counts[self.lo:self.hi], self.flags[self.lo:self.hi]

    def _slot(self, r, which=0):
        return (which * self.world + r) * self.LINE // 8

    def publish(self):
        """Announces that this rank's slice holds the results of the current round (release: x86 stores are ordered;
        the engine's wait() already made the GPU's writes visible to this thread)."""
        self._round += 1
        self._seq[self._slot(self.rank)] = self._round

    def collect(self):
        """Rank 0: blocks until every rank has published this round; returns (humans, counts, flags) of the whole
        stream (views of the shared segment, valid until the next round) and the seconds spent waiting.  Other ranks:
        returns (None, 0.0) at once."""
        if self.rank != 0:
            return None, 0.0
        t0 = time.perf_counter()
        for r in range(self.world):
            i = self._slot(r)
            while self._seq[i] < self._round:
                if time.perf_counter() - t0 > self.timeout:
                    raise TimeoutError("HostGather: rank %d did not publish round %d" % (r, self._round))
        return (self.humans, self.counts, self.flags), time.perf_counter() - t0

    def release(self):
        """Rank 0 tells the others the gathered arrays have been consumed (they may overwrite their slices)."""
        if self.rank == 0:
            self._seq[self._slot(0, 1)] = self._round

    def wait_released(self):
        """Ranks > 0: blocks until rank 0 has consumed the previous round."""
        t0 = time.perf_counter()
        i = self._slot(0, 1)
        while self._seq[i] < self._round:
            if time.perf_counter() - t0 > self.timeout:
                raise TimeoutError("HostGather: rank 0 did not release round %d" % self._round)

    def close(self):
        from . import _capi as capi
        if self._registered:
            L = capi.lib()
            for p in self._registered:
                L.opp_host_unregister(p)
            self._registered = []
        self._seq = self.humans = self.counts = self.flags = self._map = None   # unmapped when the last view dies
        if self.rank == 0:
            try:
                os.unlink(self.path)
            except OSError:
                pass


def process_stream(engine, conf, paf, rank=0, world=1, batch=None, group=None, gather=None, shard_only=False, up_buffers=None, **kw):
    """Runs this rank's shard of a stream of frames through `engine` with all its slots in flight and
    gathers on rank 0.  conf/paf hold the WHOLE stream (numpy or CUDA tensors) and only the shard is read - or, with
    shard_only=True, just this rank's shard of it (a rank need not hold the other ranks' frames).
    gather = a HostGather: results are written straight into the shared segment and rank 0 gets views of the whole
    stream; otherwise gather_results (torch.distributed) carries them.
    up_buffers = [(conf_up, paf_up), ...] device buffers for the materialised up-sampled maps, one pair per pipeline
    slot, used in turn (a batch's maps are valid from its wait() until the slot is submitted to again)."""
    from . import _capi as capi
    n = gather.n if gather is not None else (None if shard_only else int(conf.shape[0]))
    if n is None:
        raise ValueError("process_stream: shard_only needs a HostGather (it knows the stream's length)")
    lo, hi = shard_range(n, rank, world)
    base = lo if shard_only else 0
    batch = batch or engine.max_batch
    if gather is not None:
        humans, counts, flags = gather.local()
    else:
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
        if up_buffers:
            cu, pu = up_buffers[((s - lo) // batch) % len(up_buffers)]
            kw = dict(kw, conf_up=cu, paf_up=pu)
        inflight.append(engine.submit(conf[s - base:e - base], paf[s - base:e - base], out=out, **kw))
    for t in inflight:
        engine.wait(t)
    if gather is not None:
        gather.publish()
        return gather.collect()[0]
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

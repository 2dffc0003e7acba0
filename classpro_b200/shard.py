"""Read-range sharding across GPUs: contiguous ranges balanced by cumulative
compressed-profile bytes, which the FastK index gives for free (src/libfastk.c:1325-1336) and
which tracks the number of k-mers.  Reads are independent, so no rank ever needs another rank's
data: the only cross-rank step in the whole job is the ordered concatenation of the outputs
(and, in bench.py, the max-reduction of the timings)."""
import numpy as np


def shard_ranges(weights, nranks):
    """Split range(len(weights)) into nranks contiguous [beg,end) ranges of near-equal total weight.

    weights: per-read cost (compressed profile bytes or read length).  Returns a list of
    (beg, end); ranges are contiguous, ordered, cover everything, and may be empty only when
    there are fewer reads than ranks."""
    w = np.asarray(weights, dtype=np.int64)
    n = len(w)
    if nranks <= 0:
        raise ValueError("nranks must be positive")
    cum = np.concatenate([[0], np.cumsum(w)])
    total = int(cum[-1])
    cuts = [0]
    for r in range(1, nranks):
        target = total * r // nranks
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), n)
        cuts.append(c)
    cuts.append(n)
    return [(cuts[i], cuts[i + 1]) for i in range(nranks)]


def reference_thread_ranges(nreads, nthreads):
    """The reference's own split, for comparison: ceil(nreads/T) reads per thread
    (src/ClassPro.c:529-530, src/io.c:353-354)."""
    per = nreads // nthreads + (0 if nreads % nthreads == 0 else 1)
    return [(min(t * per, nreads), min((t + 1) * per, nreads)) for t in range(nthreads)]


def plan_chunk_shards(chunk_weights, nranks):
    """The shards of a read set that exists as chunks (FastK profile parts, or the chromosomes bench.py
    generates): chunk_weights = [(chunk id, per-read weights of the chunk)] in global read order.
    Returns (plans, total_reads); plans[r] = (beg, end, [(chunk id, a, b)]) -- the global read range of
    rank r and, chunk by chunk, the local read range [a,b) of the chunks it touches.  A rank only
    ever needs those chunks; borders fall inside a chunk, so neighbours may both need it."""
    first, at = [], 0
    for c, w in chunk_weights:
        first.append(at)
        at += len(w)
    allw = np.concatenate([np.asarray(w, dtype=np.int64) for c, w in chunk_weights]) if chunk_weights else np.zeros(0, np.int64)
    plans = []
    for beg, end in shard_ranges(allw, nranks):
        need = []
        for (c, w), f in zip(chunk_weights, first):
            a, b = max(beg - f, 0), min(end - f, len(w))
            if a < b:
                need.append((c, int(a), int(b)))
        plans.append((int(beg), int(end), need))
    return plans, at

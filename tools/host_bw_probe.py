#!/usr/bin/env python
"""Host <-> device bandwidth ceiling of a box (development tool, not part of the product).

    python tools/host_bw_probe.py                                   # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port 29511 tools/host_bw_probe.py                 # N GPUs at the same time

Every rank moves pinned host memory to its GPU (H2D), back (D2H), and both at once on two
streams (what the streamed remap path does: H2D of touched source rows while results travel
back), all ranks concurrently behind a barrier.  It also times the CPU side of the pageable
path: a multi-threaded memcpy (b200remap_host_pack_runs) of the same volume.  Rank 0 prints one
JSON object: per-rank and aggregate GB/s -- the ceiling the end-to-end slices/s of bench.py
must be read against (one C3 slice = 219 MB in + 193 MB out).
"""
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local)
    dev = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('gloo')
    nbytes = 1 << 30
    h_in = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_out = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_in.fill_(1)
    d_in = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_out = torch.ones(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def barrier():
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()

    def timed(fn, reps=5):
        fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        torch.cuda.synchronize(dev)
        dt = time.perf_counter() - t0
        barrier()
        return reps / dt

    def h2d():
        with torch.cuda.stream(s1):
            d_in.copy_(h_in, non_blocking=True)

    def d2h():
        with torch.cuda.stream(s2):
            h_out.copy_(d_out, non_blocking=True)

    def both():
        h2d()
        d2h()

    res = {'rank': rank,
           'h2d_GBps': timed(h2d) * nbytes / 1e9,
           'd2h_GBps': timed(d2h) * nbytes / 1e9,
           'duplex_GBps_sum': timed(both) * 2 * nbytes / 1e9}
    # CPU side of the pageable path: threads copying pageable -> pinned
    try:
        from pyremap_b200 import _cabi
        page = np.ones(nbytes, dtype=np.uint8)
        off = np.arange(0, nbytes, 1 << 20, dtype=np.int64)
        ln = np.full(off.size, 1 << 20, dtype=np.int64)
        cores = len(os.sched_getaffinity(0))
        for threads in sorted({1, 4, 8, min(16, cores), min(32, cores)}):
            def pack():
                _cabi.host_pack_runs(page.ctypes.data, h_in.data_ptr(), off, off, ln, threads)
            res[f'host_memcpy_{threads}thr_GBps'] = timed(pack, reps=3) * nbytes / 1e9
        res['cores_allowed'] = cores
    except Exception as exc:      # noqa: BLE001
        res['host_memcpy'] = f'unavailable: {exc}'
    try:
        res['thp'] = open('/sys/kernel/mm/transparent_hugepage/enabled').read().strip()
    except OSError:
        res['thp'] = 'unknown'
    allr = [res]
    if world > 1:
        allr = [None] * world
        dist.all_gather_object(allr, res)
    if rank == 0:
        agg = {k: sum(r[k] for r in allr) for k in ('h2d_GBps', 'd2h_GBps', 'duplex_GBps_sum')}
        print(json.dumps({'n_gpus': world, 'aggregate': agg, 'per_rank': allr,
                          'note': 'all ranks concurrently; 1 GiB pinned buffers; duplex = H2D and D2H '
                                  'on two streams at once, sum of both directions'}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()

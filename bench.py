#!/usr/bin/env python
"""Benchmark of the weight-application hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (rank 0 only)

Workload (BASELINE.json configs[4], the multi-slice form of configs[2]): daily
slices of a synthetic oRRS18to6-size MPAS mesh -- 3 693 225 cells x 80 levels, fp64,
bathymetry-masked -- remapped to the 601 x 501 Antarctic stereographic grid with
renormalisation threshold 0.01 (the reference's masked branch,
pyremap/remapper/remap_numpy.py:263-266).  One STEP = one sweep of
``--slices`` (default 365) slices per GPU, issued as (Time=8, nCells, nVertLevels)
batches, i.e. one fused launch per 8 slices, cycling over a ring of 8 distinct
2.36 GB slices resident in HBM (so consecutive launches never share cache lines of X;
the ring is 19 GB >> 126 MB L2).  Weak scaling: every rank sweeps its own slices,
weights replicated, no collective in the data path.

One JSON line is printed by rank 0; see README/DESIGN.md for the keys.
"""

from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_LEVELS = 80
RING = int(os.environ.get('B200REMAP_BENCH_RING', '8'))     # distinct slices resident in HBM
BATCH = int(os.environ.get('B200REMAP_BENCH_BATCH', '8'))   # slices per fused launch (Time chunk)
THRESHOLD = 0.01
METRIC = 'remap_field_slices_per_s'
UNIT = 'field-slices/s'


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=10)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--slices', type=int, default=365, help='slices per GPU per step')
    ap.add_argument('--scale', type=float, default=1.0, help='mesh scale (1 = BASELINE size)')
    ap.add_argument('--mode', default='masked', choices=['masked', 'unmasked'])
    ap.add_argument('--no-e2e', action='store_true')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--e2e-slices', type=int, default=8, help='slices per e2e call')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    return ap.parse_args()


def workload_name(args, m):
    return (f'C5/C3: {args.slices} daily slices per GPU of synthetic MPAS-like {m.n_a}-cell x '
            f'{N_LEVELS}-level fp64 fields -> {m.dst_descriptor.dim_sizes[1]}x'
            f'{m.dst_descriptor.dim_sizes[0]} Antarctic stereographic, {args.mode} branch'
            + (f' thr={THRESHOLD}' if args.mode == 'masked' else ''))


def algorithmic_bytes(m, csr_info, K, w_in=8, with_fracb=False):
    """B = nnz*12 + (n_b+1)*4 + n_touched*K*w_in + n_b*K*8 [+ n_b*8]."""
    nnz, n_b, n_touched = csr_info['nnz'], csr_info['n_row'], csr_info['n_touched']
    b = nnz * 12 + (n_b + 1) * 4 + n_touched * K * w_in + n_b * K * 8
    if with_fracb:
        b += n_b * 8
    b_full = nnz * 12 + (n_b + 1) * 4 + m.n_a * K * w_in + n_b * K * 8
    return b, b_full


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    FIELDS = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
              'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
              'clocks_event_reasons.sw_power_cap')

    def __init__(self, device_index):
        self.samples = []
        self.proc = None
        self.t0 = self.t1 = None
        try:
            uuid = None
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = str(device_index if not vis else vis.split(',')[device_index])
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '50', '-i', idx], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
            del uuid
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.samples.append((time.time(), line.strip()))

    def window(self, t0, t1):
        self.t0, self.t1 = t0, t1

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        rows = [s for t, s in self.samples if self.t0 is None or self.t0 - 0.05 <= t <= self.t1 + 0.05]
        if not rows:
            rows = [s for _, s in self.samples[-3:]]
        sm, smax, power = [], [], []
        reasons = set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for r in rows:
            parts = [p.strip() for p in r.split(',')]
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
                power.append(float(parts[2]))
            except (ValueError, IndexError):
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith('active'):
                    reasons.add(name)
        return {'sm_mhz': float(np.median(sm)) if sm else None,
                'sm_max_mhz': max(smax) if smax else None,
                'power_w_max': max(power) if power else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ----------------------------------------------------------------------------
# CPU reference arm / baseline (the ONLY place bench.py executes oracle/)
# ----------------------------------------------------------------------------
def cpu_reference_slice_rate(m, mode, seconds_budget, steps=None, warmup=0):
    """Time the reference's CPU algorithm (oracle/remap_oracle.remap_array_stepwise:
    scipy csr.dot + the reference's NumPy passes, 1 thread like the reference) on a
    bounded sample: one slice restricted to as many of the 80 levels as fit the budget."""
    from oracle import remap_oracle
    from pyremap_b200 import synthetic as syn
    A = remap_oracle.build_matrix(m.S, m.row, m.col, m.n_b, m.n_a)
    lv = syn.bathymetry_levels(m.n_a, N_LEVELS, seed=5)

    def make(levels):
        return syn.ocean_field(m.n_a, levels, seed=6,
                               max_level=lv if mode == 'masked' else None)

    def one(raw):
        t = time.perf_counter()
        # remap_numpy.py:201-204: the any-NaN test that selects the branch
        field = raw
        nanmask = np.isnan(raw)
        if np.count_nonzero(nanmask) > 0:
            field = np.ma.masked_array(raw, nanmask)
        out = remap_oracle.remap_array_stepwise(A, m.frac_b, m.dst_grid_dims, field, [0],
                                                THRESHOLD if mode == 'masked' else None)
        dt = time.perf_counter() - t
        assert out.shape[-1] == raw.shape[-1]
        return dt

    # calibrate on 8 levels, then pick the level count that fits the budget
    probe = make(8)
    one(probe)
    t8 = one(probe)
    n_runs = (steps + warmup) if steps else 3
    per_run = seconds_budget / max(1, n_runs)
    levels = int(max(8, min(N_LEVELS, 8 * per_run / max(t8, 1e-6))))
    levels -= levels % 8
    field = probe if levels == 8 else make(levels)
    for _ in range(warmup):
        one(field)
    times = [one(field) for _ in range(steps if steps else 3)]
    dt = float(np.mean(times)) if steps else float(min(times))
    rate = (levels / N_LEVELS) / dt
    sample = (f'one slice restricted to {levels}/{N_LEVELS} levels ({m.n_a} cells), '
              f'{len(times)} run(s), {dt:.3f} s each; isnan scan + scipy csr.dot + NumPy passes of '
              f'remap_numpy.py:201-204,256-278, single thread (scipy/numpy use 1)')
    return rate, sample, dt, levels


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    from pyremap_b200 import synthetic as syn
    m = syn.make_c3(scale=args.scale)
    t0 = time.time()
    rate, sample, dt, levels = cpu_reference_slice_rate(m, args.mode, 150.0, steps=args.steps,
                                                        warmup=args.warmup)
    out = {
        'impl': 'reference', 'metric': METRIC, 'value': rate, 'unit': UNIT,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': dt * 1e3, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': workload_name(args, m), 'mesh_scale': args.scale,
                   'step': 'one bounded sample (see cpu_baseline.sample)'},
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores_available': os.cpu_count()},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.time() - t0,
    }
    print(json.dumps(out), file=JSON_OUT, flush=True)


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def make_ring(torch, m, device, mode, seed):
    """[RING, nCells, 80] fp64 distinct slices generated on the device."""
    from pyremap_b200 import synthetic as syn
    g = torch.Generator(device=device).manual_seed(seed)
    ring = torch.empty((RING, m.n_a, N_LEVELS), dtype=torch.float64, device=device)
    for s in range(RING):
        ring[s].uniform_(-2.0, 30.0, generator=g)
    if mode == 'masked':
        lv = torch.from_numpy(syn.bathymetry_levels(m.n_a, N_LEVELS, seed=5)).to(device)
        dry = torch.arange(N_LEVELS, device=device)[None, :] >= lv[:, None]
        ring.masked_fill_(dry[None], float('nan'))
        del dry
    return ring


def run_b200(args):
    import torch
    import torch.distributed as dist

    from pyremap_b200 import _cabi, mapfile, synthetic as syn
    import pyremap_b200

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    m = syn.make_c3(scale=args.scale)
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                   m.n_b, m.n_a)
    matrix = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
    csr = matrix.on_device(local)
    info = {'n_row': csr.n_row, 'n_col': csr.n_col, 'nnz': csr.nnz, 'n_touched': csr.n_touched,
            'max_row_nnz': csr.max_row_nnz, 'empty_rows': csr.n_empty_rows}
    mode_code = _cabi.MODE_MASKED if args.mode == 'masked' else _cabi.MODE_FRACB
    b_slice, b_slice_full = algorithmic_bytes(m, info, N_LEVELS, 8, args.mode != 'masked')

    ring = make_ring(torch, m, device, args.mode, seed=100 + rank)
    y = torch.empty((BATCH, m.n_b, N_LEVELS), dtype=torch.float64, device=device)
    stream = torch.cuda.current_stream(device)

    # the launches of one step: full batches plus one ragged batch
    plan = []
    left = args.slices
    while left > 0:
        nb = min(BATCH, left)
        plan.append(nb)
        left -= nb

    def sweep():
        for nb in plan:
            csr.spmm(ring.data_ptr(), _cabi.F64, N_LEVELS, N_LEVELS, nb, m.n_a * N_LEVELS,
                     y.data_ptr(), N_LEVELS, m.n_b * N_LEVELS, mode_code, THRESHOLD,
                     stream=stream.cuda_stream)
        return len(plan)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    for _ in range(max(args.warmup, 3)):
        sweep()
    barrier()

    sampler = ClockSampler(local) if rank == 0 else None
    # launch-level timing of the dominant kernel (full 8-slice batches) inside the region
    ev_l0 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev_l1 = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps)]
    ev0 = torch.cuda.Event(enable_timing=True)
    ev1 = torch.cuda.Event(enable_timing=True)
    launches = 0
    barrier()
    t_wall0 = time.time()
    ev0.record(stream)
    for s in range(args.steps):
        ev_l0[s].record(stream)
        launches += sweep()
        ev_l1[s].record(stream)
    ev1.record(stream)
    barrier()
    t_wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    if sampler:
        sampler.window(t_wall0, t_wall1)
        clocks = sampler.stop()
    t = torch.tensor([elapsed_ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    elapsed_ms = float(t.item())
    total_slices = world * args.steps * args.slices
    value = total_slices / (elapsed_ms * 1e-3)

    # dominant kernel: average duration of a full-batch launch = step time / slices * BATCH
    step_ms = [a.elapsed_time(b) for a, b in zip(ev_l0, ev_l1)]
    launch_ms = float(np.mean(step_ms)) / args.slices * BATCH
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        peak, peak_src = float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        peak, peak_src = 6650.0, 'fallback (B200_PROFILING.md)'
    achieved = b_slice * BATCH / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed capture
        cap = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if args.mode == 'masked' and args.scale == 1.0:
            traffic, traffic_src = cap['traffic_bytes_per_launch'], cap['source']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': peak_src,
                # AUTO (b200remap_spmm): batched masked sweeps -> wrow_kernel, else pbin_kernel
                'kernel': ('wrow_kernel' if args.mode == 'masked' and BATCH >= 2 else 'pbin_kernel')
                + '<double,VEC=4,MODE=%d>' % mode_code,
                'launch_ms': launch_ms, 'algorithmic_bytes_per_launch': b_slice * BATCH,
                'algorithmic_bytes_per_slice': b_slice,
                'full_x_bytes_per_slice': b_slice_full,
                'frac_of_nominal_8TBs': achieved / 8000.0,
                'nnz_col_per_s': info['nnz'] * N_LEVELS * BATCH / (launch_ms * 1e-3)}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic',
        'config': {'workload': workload_name(args, m), 'mesh_scale': args.scale,
                   'slices_per_gpu_per_step': args.slices, 'slices_per_launch': BATCH,
                   'ring_slices': RING, 'levels': N_LEVELS, 'map': info,
                   'l2_policy': 'inputs larger than L2: each launch reads 8 distinct slices '
                                f'({RING * m.n_a * N_LEVELS * 8 / 1e9:.1f} GB ring); no flush needed',
                   'parallelism': f'K-sharded replicas x{world}, no collective'},
        'aggregate_algorithmic_GBps': value * b_slice / 1e9,
        'roofline': roofline, 'gpu_launches': launches,
    }

    # -------- end to end through the public API with host buffers --------
    if not args.no_e2e:
        out['e2e'] = measure_e2e(args, torch, dist, m, matrix, device, world, ring)
    if rank == 0:
        out['clocks'] = clocks
        if world == 1 and not args.no_cpu_baseline:
            rate, sample, _, _ = cpu_reference_slice_rate(m, args.mode, args.cpu_seconds)
            out['cpu_baseline'] = {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                                   'sample': sample, 'host_cores_available': os.cpu_count()}
        print(json.dumps(out), file=JSON_OUT, flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_e2e(args, torch, dist, m, matrix, device, world, ring):
    """Slices/s through ``Remapper.remap_array`` with HOST buffers: a pinned
    ``(Time, nCells, nVertLevels)`` ndarray in, an ndarray out; inside the call every slice is
    copied host->device (only the source rows the map touches), remapped, and copied back."""
    import pyremap_b200
    r = pyremap_b200.Remapper(map_filename='in-memory', src_descriptor=m.src_descriptor,
                              dst_descriptor=m.dst_descriptor)
    r._matrix = matrix
    r._ds_map = mapfile_dataset(m)
    r.device = device.index
    # every rank pins T input slices (2.36 GB each) plus its results: keep the host footprint of
    # an 8-rank run moderate (4 slices per call there, 8 on 1-2 GPUs)
    T = max(1, min(args.e2e_slices if world <= 2 else min(args.e2e_slices, 4), RING))
    host_t = torch.empty((T, m.n_a, N_LEVELS), dtype=torch.float64, pin_memory=True)
    host_t.copy_(ring[:T])                      # same synthetic slices, now in host memory
    torch.cuda.synchronize(device)
    host = host_t.numpy()
    thr = THRESHOLD if args.mode == 'masked' else None
    # the result goes to a preallocated pinned block (out=): the device->host copies then land in
    # the caller's memory directly; without out= the call returns a fresh pageable array filled
    # through a pinned staging ring by CPU threads
    out = torch.empty((T,) + tuple(m.dst_descriptor.dim_sizes) + (N_LEVELS,), dtype=torch.float64,
                      pin_memory=True).numpy()
    for _ in range(2):                          # warm-up: cover CSR, streams
        r.remap_array(host, [1], thr, out=out)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize(device)
    calls = 3
    t0 = time.perf_counter()
    for _ in range(calls):
        r.remap_array(host, [1], thr, out=out)
    torch.cuda.synchronize(device)
    dt = time.perf_counter() - t0
    tt = torch.tensor([dt], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dt = float(tt.item())
    assert out.shape == (T, m.dst_descriptor.dim_sizes[0], m.dst_descriptor.dim_sizes[1], N_LEVELS)
    # the same call without out=: a fresh pageable result per call (what the xarray path returns)
    fresh = r.remap_array(host, [1], thr)
    torch.cuda.synchronize(device)
    t1 = time.perf_counter()
    fresh = r.remap_array(host, [1], thr)
    torch.cuda.synchronize(device)
    dt_fresh = time.perf_counter() - t1
    assert fresh.shape == out.shape
    cov = matrix.cover_exact()                  # pinned input: exactly the touched rows travel
    rows_copied = cov['n_cover'] if cov else m.n_a
    n_runs = int(cov['run_start'].size) if cov else 1
    return {'value': world * calls * T / dt, 'unit': UNIT,
            'h2d_bytes_per_step': int(T * rows_copied * N_LEVELS * 8),
            'd2h_bytes_per_step': int(T * m.n_b * N_LEVELS * 8),
            'step': f'one call of Remapper.remap_array(pinned host ndarray (Time={T}, nCells, '
                    f'nVertLevels), out=pinned host ndarray); {rows_copied} of {m.n_a} source rows copied per '
                    f'slice (the {cov["n_touched"] if cov else m.n_a} rows the map touches plus bridged '
                    f'gaps of <= {cov["bridged_gap"] if cov else 0} rows: {n_runs} contiguous runs, one '
                    f'batched DMA submission per slice, full duplex with the D2H of results)',
            'calls_timed': calls, 'slices_per_call': T, 'ms_per_slice': dt / (calls * T) * 1e3,
            'fresh_pageable_result_ms_per_slice_this_rank': dt_fresh / T * 1e3,
            'host_nan_scan': 'whole variable, native early-exit scan (branch selection)'}


def mapfile_dataset(m):
    from pyremap_b200 import mapfile
    return mapfile.MapDataset({'dst_grid_dims': m.dst_grid_dims, 'frac_b': m.frac_b,
                               'src_grid_dims': m.src_grid_dims}, {})


JSON_OUT = sys.stdout


def claim_stdout():
    """Keep stdout for the one JSON line: everything else that writes to file descriptor 1
    (NCCL prints its version banner there when NCCL_DEBUG is set) goes to stderr."""
    global JSON_OUT
    sys.stdout.flush()
    JSON_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)


def main():
    args = parse_args()
    claim_stdout()
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_b200(args)


if __name__ == '__main__':
    main()

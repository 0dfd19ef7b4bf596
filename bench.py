#!/usr/bin/env python
"""Benchmark of the weight-application hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one process per GPU)
    python bench.py --impl reference --steps K --warmup W    # CPU reference arm (rank 0 only)

Headline workload (BASELINE.json configs[4], the multi-slice form of configs[2]): daily
slices of a synthetic oRRS18to6-size MPAS mesh -- 3 693 225 cells x 80 levels, fp64,
bathymetry-masked -- remapped to the 601 x 501 Antarctic stereographic grid with
renormalisation threshold 0.01 (the reference's masked branch,
pyremap/remapper/remap_numpy.py:263-266).  One STEP = one sweep of ``--slices`` (default 365)
slices per GPU, issued as (Time=8, nCells, nVertLevels) batches, i.e. one fused launch per 8
slices, cycling over a ring of 8 distinct 2.36 GB slices resident in HBM (18.9 GB >> 126 MB L2).
Weak scaling: every rank sweeps its own slices, weights replicated, no collective in the data
path.

One JSON line is printed by rank 0.  Besides the contract keys it carries
  roofline      dominant kernel: algorithmic bytes per launch (weights counted ONCE per launch)
                / mean launch time inside the timed region, against MEASURED_PEAKS.json
  sharded       strong scaling through the product's ShardedRemap: 365 slices in total dealt to
                the ranks, optional NCCL gather timed, sharded == unsharded checked bit for bit
  e2e           Remapper.remap_array with pinned host buffers (H2D + D2H in the timed region)
  e2e_dropin    Remapper.remap_numpy(Dataset) with pageable arrays in, fresh arrays out
  configs       every other BASELINE config timed in-process with CUDA events (N = 1 only)
  cpu_baseline  the reference's CPU algorithm on a bounded sample (N = 1 only)
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
    ap.add_argument('--no-configs', action='store_true')
    ap.add_argument('--no-sharded', action='store_true')
    ap.add_argument('--e2e-slices', type=int, default=8, help='slices per e2e call')
    ap.add_argument('--cpu-seconds', type=float, default=20.0)
    ap.add_argument('--order', default='rings', choices=['rings', 'blocks', 'shuffle'],
                    help='numbering of the synthetic source cells (mesh-order sensitivity runs)')
    return ap.parse_args()


def workload_name(args):
    n_a = max(2000, int(round(3693225 * args.scale)))
    lin = 1.0 / np.sqrt(args.scale)
    nx, ny = int(6000.0 / (10.0 * lin)) + 1, int(5000.0 / (10.0 * lin)) + 1
    return (f'C5/C3: {args.slices} daily slices per GPU of synthetic MPAS-like {n_a}-cell x '
            f'{N_LEVELS}-level fp64 fields -> {nx}x{ny} Antarctic stereographic, {args.mode} branch'
            + (f' thr={THRESHOLD}' if args.mode == 'masked' else ''))


def bench_config(args):
    """The ``config`` object: a function of the command line only, so that both arms (ours and
    ``--impl reference``) print the identical dict."""
    return {'workload': workload_name(args), 'mesh_scale': args.scale,
            'slices_per_gpu_per_step': args.slices, 'levels': N_LEVELS, 'dtype': 'f64',
            'branch': args.mode, 'threshold': THRESHOLD if args.mode == 'masked' else None,
            'gpu_batching': f'(Time={BATCH}, nCells, nVertLevels) per fused launch over a ring of '
                            f'{RING} distinct slices resident in HBM (larger than L2: no flush needed)',
            'parallelism': 'weights replicated, slices sharded over the ranks, no collective',
            'source_cell_order': args.order}


def weight_bytes(info, with_fracb):
    return info['nnz'] * 12 + (info['n_row'] + 1) * 4 + (info['n_row'] * 8 if with_fracb else 0)


def field_bytes(info, K, w_in=8, w_out=8):
    return info['n_touched'] * K * w_in + info['n_row'] * K * w_out


def launch_bytes(info, K, nb, with_fracb=False, w_in=8, w_out=8):
    """SURVEY.md section 8(d) per launch: the weights once, the fields of its nb slices:
    B = nnz*12 + (n_b+1)*4 [+ n_b*8] + nb * (n_touched*K*w_in + n_b*K*w_out)."""
    return weight_bytes(info, with_fracb) + nb * field_bytes(info, K, w_in, w_out)


def hbm_peak():
    try:
        peaks = json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))
        return float(peaks['hbm_gbs']), 'measured (MEASURED_PEAKS.json hbm_gbs)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


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
            vis = os.environ.get('CUDA_VISIBLE_DEVICES')
            idx = str(device_index if not vis else vis.split(',')[device_index])
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--query-gpu={self.FIELDS}', '--format=csv,noheader,nounits',
                 '-lms', '50', '-i', idx], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL,
                text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
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
        'config': bench_config(args),
        'step': 'one bounded sample of the workload (see cpu_baseline.sample)',
        'cpu_baseline': {'value': rate, 'unit': UNIT, 'cores': 1, 'kind': 'port',
                         'sample': sample, 'host_cores_available': os.cpu_count()},
        'e2e': {'value': rate, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0, 'wall_s': time.time() - t0,
    }
    print(json.dumps(out), file=JSON_OUT, flush=True)


# ----------------------------------------------------------------------------
# helpers of our arm
# ----------------------------------------------------------------------------
def bind_to_gpu_numa_node(local):
    """Run this rank's threads (and first-touch its pinned buffers) on the NUMA node the GPU
    hangs off.  Returns a short description for the JSON line."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(local)
        path = (f'/sys/bus/pci/devices/{prop.pci_domain_id:04x}:{prop.pci_bus_id:02x}:'
                f'{prop.pci_device_id:02x}.0/numa_node')
        node = int(open(path).read().strip())
        if node < 0:
            return 'gpu numa node unknown (-1): not bound'
        cpus = open(f'/sys/devices/system/node/node{node}/cpulist').read().strip()
        ids = []
        for part in cpus.split(','):
            a, _, b = part.partition('-')
            ids.extend(range(int(a), int(b or a) + 1))
        allowed = sorted(set(ids) & set(os.sched_getaffinity(0)))
        if allowed:
            os.sched_setaffinity(0, allowed)
        n_nodes = len([d for d in os.listdir('/sys/devices/system/node') if d.startswith('node')])
        return f'numa node {node} of {n_nodes} ({len(allowed)} cpus)'
    except Exception as exc:      # noqa: BLE001 - best effort, reported
        return f'not bound ({type(exc).__name__})'


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


def csr_of(m, device_index):
    from pyremap_b200 import mapfile
    ip, ix, d = mapfile.coo_to_csr(m.S, m.row.astype(np.int64) - 1, m.col.astype(np.int64) - 1,
                                   m.n_b, m.n_a)
    matrix = mapfile.WeightMatrix(ip, ix, d, (m.n_b, m.n_a), m.frac_b)
    csr = matrix.on_device(device_index)
    info = {'n_row': csr.n_row, 'n_col': csr.n_col, 'nnz': csr.nnz, 'n_touched': csr.n_touched,
            'max_row_nnz': csr.max_row_nnz, 'empty_rows': csr.n_empty_rows}
    return matrix, csr, info


def kernel_name(csr, dtype='double', vec=4, mode=2, K=N_LEVELS):
    from pyremap_b200 import _cabi
    code = _cabi.F64 if dtype == 'double' else _cabi.F32
    return f'{_cabi.KERNEL_NAMES.get(csr.auto_kernel(code, K), "?")}<{dtype},VEC={vec},MODE={mode}>'


def time_launches(torch, fn, reps=20, warm=3, flush=None):
    """Median / best duration (ms) of ``fn(i)`` by CUDA events on the current stream; ``flush``
    (a device buffer larger than L2) is overwritten between iterations when the inputs of
    ``fn`` would otherwise stay cache-resident."""
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
           for _ in range(reps)]
    for i, (a, b) in enumerate(evs):
        if flush is not None:
            flush.zero_()
        a.record()
        fn(i)
        b.record()
    torch.cuda.synchronize()
    ts = sorted(a.elapsed_time(b) for a, b in evs)
    return ts[len(ts) // 2], ts[0]


# ----------------------------------------------------------------------------
# our arm
# ----------------------------------------------------------------------------
def run_b200(args):
    import torch
    import torch.distributed as dist

    from pyremap_b200 import _cabi, synthetic as syn

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py: no CUDA device; the product has no CPU path')
    torch.cuda.set_device(local)
    device = torch.device('cuda', local)
    numa = bind_to_gpu_numa_node(local)
    if world > 1:
        dist.init_process_group('nccl', device_id=device)

    m = syn.make_c3(scale=args.scale, order=args.order)
    matrix, csr, info = csr_of(m, local)
    masked = args.mode == 'masked'
    mode_code = _cabi.MODE_MASKED if masked else _cabi.MODE_FRACB
    b_launch = launch_bytes(info, N_LEVELS, BATCH, with_fracb=not masked)
    b_slice = launch_bytes(info, N_LEVELS, 1, with_fracb=not masked)

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
    clocks = None
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
    peak, peak_src = hbm_peak()
    achieved = b_launch / (launch_ms * 1e-3) / 1e9
    traffic, traffic_src = None, None
    try:       # dram__bytes_read.sum + dram__bytes_write.sum of one launch, from the committed capture
        cap = json.load(open(os.path.join(ROOT, 'profiles', 'ncu_traffic.json')))
        if masked and args.scale == 1.0:
            traffic, traffic_src = cap['traffic_bytes_per_launch'], cap['source']
    except Exception:
        pass
    roofline = {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                'frac': achieved / peak, 'traffic': traffic, 'traffic_source': traffic_src,
                'peak_source': peak_src, 'kernel': kernel_name(csr, mode=mode_code),
                'launch_ms': launch_ms, 'algorithmic_bytes_per_launch': b_launch,
                'algorithmic_bytes_note': 'weights counted once per launch + 8 x (touched source '
                                          'rows + result rows), SURVEY 8d',
                'algorithmic_bytes_per_slice_alone': b_slice,
                'full_x_bytes_per_launch': weight_bytes(info, not masked) + BATCH * (
                    m.n_a * N_LEVELS * 8 + m.n_b * N_LEVELS * 8),
                'frac_of_nominal_8TBs': achieved / 8000.0,
                'nnz_col_per_s': info['nnz'] * N_LEVELS * BATCH / (launch_ms * 1e-3)}

    out = {
        'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps,
        'warmup': max(args.warmup, 3), 'ms_per_step': elapsed_ms / args.steps,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
        'data': 'synthetic', 'config': bench_config(args), 'map': info, 'host_binding': numa,
        'aggregate_algorithmic_GBps': value * b_launch / BATCH / 1e9,
        'roofline': roofline, 'gpu_launches': launches,
    }

    # -------- strong scaling through the product's sharding driver --------
    if not args.no_sharded:
        out['sharded'] = measure_sharded(args, torch, dist, m, matrix, device, world, ring)
    # -------- end to end through the public API with host buffers --------
    if not args.no_e2e:
        out['e2e'] = measure_e2e(args, torch, dist, m, matrix, device, world, ring)
        out['e2e_dropin'] = measure_dropin(args, torch, dist, m, matrix, device, world, ring)
    if world == 1 and not args.no_configs and args.scale == 1.0 and args.order == 'rings':
        del y
        out['configs'] = measure_configs(args, torch, device, m, csr, info, ring, peak)
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


def make_remapper(m, matrix, device):
    import pyremap_b200
    r = pyremap_b200.Remapper(map_filename='in-memory', src_descriptor=m.src_descriptor,
                              dst_descriptor=m.dst_descriptor)
    r._matrix = matrix
    r._ds_map = mapfile_dataset(m)
    r.device = device.index
    return r


def measure_sharded(args, torch, dist, m, matrix, device, world, ring):
    """BASELINE configs[4] as a STRONG-scaling sweep through ``pyremap_b200.sharding``:
    ``--slices`` (365) slices in total are dealt to the ranks in contiguous blocks
    (``ShardedRemap.sweep``, chunks of 8 slices per fused launch, weights replicated, no
    collective in the data path); the optional NCCL all-gather of one 8-slice result is timed
    separately, and one batch is checked bit for bit against the unsharded result."""
    from pyremap_b200.sharding import ShardedRemap
    thr = THRESHOLD if args.mode == 'masked' else None
    r = make_remapper(m, matrix, device)
    sh = ShardedRemap(r, [1], thr)
    n_total = args.slices
    masked = args.mode == 'masked'

    def loader(a, b):                      # slices [a, b) of the synthetic year: a ring view
        return ring[:b - a]

    count = [0]

    def sink(a, b, out):
        count[0] += out.shape[0]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    sh.sweep(loader, n_total, chunk=BATCH, sink=sink, masked=masked)     # warm-up
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 3
    count[0] = 0
    ev0.record()
    for _ in range(reps):
        sh.sweep(loader, n_total, chunk=BATCH, sink=sink, masked=masked)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1) / reps
    t = torch.tensor([ms], dtype=torch.float64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms = float(t.item())
    lo, hi = sh.local_slices(n_total)
    assert count[0] == reps * (hi - lo)
    res = {'slices_total': n_total, 'scaling': 'strong', 'ms_per_sweep': ms,
           'value': n_total / (ms * 1e-3), 'unit': UNIT,
           'api': 'pyremap_b200.sharding.ShardedRemap.sweep -> Remapper.remap_array(CUDA tensor)',
           'slices_this_rank': hi - lo}
    # one batch, sharded vs unsharded, bit for bit; and the optional gather.  Every rank holds
    # its own synthetic ring, so rank 0's batch is broadcast first (NCCL over NVLink)
    T = min(BATCH, RING)
    if world > 1:
        dist.broadcast(ring[:T], src=0)
    local = sh.remap_local(ring[:T])
    torch.cuda.synchronize(device)
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    full = sh.gather(local, T)                                        # warm-up (NCCL setup)
    barrier()
    g0.record()
    full = sh.gather(local, T)
    g1.record()
    torch.cuda.synchronize(device)
    whole = r.remap_array(ring[:T], [1], thr, return_torch=True)
    same = bool(torch.equal(torch.isnan(full), torch.isnan(whole)) and
                torch.equal(torch.nan_to_num(full).view(torch.int64),
                            torch.nan_to_num(whole).view(torch.int64)))
    flag = torch.tensor([1 if same else 0], dtype=torch.int32, device=device)
    if world > 1:
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    res['sharded_parity'] = bool(int(flag.item()))
    res['gather_ms_per_8_slices'] = g0.elapsed_time(g1) if world > 1 else 0.0
    res['gather_bytes'] = int(T * m.n_b * N_LEVELS * 8)
    del full, whole, local
    return res


def measure_e2e(args, torch, dist, m, matrix, device, world, ring):
    """Slices/s through ``Remapper.remap_array`` with HOST buffers: a pinned
    ``(Time, nCells, nVertLevels)`` ndarray in, an ndarray out; inside the call every slice is
    copied host->device (only the source rows the map touches), remapped, and copied back."""
    r = make_remapper(m, matrix, device)
    T = max(1, min(args.e2e_slices, RING))          # the same call shape at every rank count
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
    cov = matrix.cover_exact()                  # pinned input: exactly the touched rows travel
    rows_copied = cov['n_cover'] if cov else m.n_a
    n_runs = int(cov['run_start'].size) if cov else 1
    res = {'value': world * calls * T / dt, 'unit': UNIT,
           'h2d_bytes_per_step': int(T * rows_copied * N_LEVELS * 8),
           'd2h_bytes_per_step': int(T * m.n_b * N_LEVELS * 8),
           'step': f'one call of Remapper.remap_array(pinned host ndarray (Time={T}, nCells, '
                   f'nVertLevels), out=pinned host ndarray); {rows_copied} of {m.n_a} source rows copied per '
                   f'slice (the {cov["n_touched"] if cov else m.n_a} rows the map touches plus bridged '
                   f'gaps of <= {cov["bridged_gap"] if cov else 0} rows: {n_runs} contiguous runs, one '
                   f'batched DMA submission per slice, full duplex with the D2H of results)',
           'calls_timed': calls, 'slices_per_call': T, 'ms_per_slice': dt / (calls * T) * 1e3,
           'host_nan_scan': 'whole variable, native early-exit scan (branch selection)'}
    del host_t, out
    return res


def measure_dropin(args, torch, dist, m, matrix, device, world, ring):
    """The drop-in call a pyremap user makes: ``Remapper.remap_numpy(ds, thr)`` on a Dataset of
    PAGEABLE arrays (what ``da.values`` is), fresh result arrays out.  Two 3-D variables
    ``(Time=4, nCells, nVertLevels=80)`` plus a passthrough variable; all variables with the
    source dims share one streamed pipeline."""
    try:
        import xarray as xr
        if not hasattr(xr, 'Dataset'):
            raise ImportError
        flavour = 'xarray'
    except ImportError:
        sys.path.insert(0, os.path.join(ROOT, 'tests'))
        import minixarray as xr            # the repo's small stand-in (xarray is not in this image)
        sys.modules['xarray'] = xr
        flavour = 'tests/minixarray.py (xarray is not installed in this image)'
    r = make_remapper(m, matrix, device)
    T = max(1, min(4, RING // 2))
    a = ring[:T].cpu().numpy()                       # pageable
    b = ring[T:2 * T].cpu().numpy()
    ds = xr.Dataset(
        {'temperature': (('Time', 'nCells', 'nVertLevels'), a),
         'salinity': (('Time', 'nCells', 'nVertLevels'), b),
         'xtime': (('Time', 'StrLen'), np.zeros((T, 4), dtype='S1'))},
        coords={'Time': np.arange(T, dtype=np.float64)})
    thr = THRESHOLD if args.mode == 'masked' else None
    from pyremap_b200 import engine

    def timed_call():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)
        t0 = time.perf_counter()
        res = r.remap_numpy(ds, thr)
        torch.cuda.synchronize(device)
        dt_ = time.perf_counter() - t0
        tt = torch.tensor([dt_], dtype=torch.float64, device=device)
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        return res, float(tt.item())

    # steady state of a loop `out = remapper.remap_numpy(ds_t, thr)`: the previous result has been
    # dropped, its buffers were page-locked by the pool's helper thread and are written directly
    out = None
    for _ in range(2):                               # warm-up: streams, staging rings, result pool
        out = None
        engine._RESULTS.wait_idle()
        out, _ = timed_call()
    out = None
    engine._RESULTS.wait_idle()
    out, dt = timed_call()
    # the same call while the caller still holds the previous result (no buffer to recycle:
    # pageable result filled through the pinned ring by CPU threads)
    keep = out
    out, dt_retained = timed_call()
    del keep
    n = 2 * T
    assert out['temperature'].values.shape == (T,) + tuple(m.dst_descriptor.dim_sizes) + (N_LEVELS,)
    cov = matrix.cover_exact()
    rows_copied = cov['n_cover'] if cov else m.n_a
    res = {'value': world * n / dt, 'unit': UNIT, 'ms_per_slice': dt / n * 1e3,
           'h2d_bytes_per_step': int(n * rows_copied * N_LEVELS * 8),
           'd2h_bytes_per_step': int(n * m.n_b * N_LEVELS * 8),
           'step': f'one call of Remapper.remap_numpy(Dataset{{temperature, salinity: (Time={T}, nCells, '
                   f'nVertLevels={N_LEVELS}) float64, xtime}}, {thr}): pageable arrays in, fresh '
                   'float64 arrays out; host NaN scan per variable, CPU threads pack the touched runs '
                   'into pinned staging, one pipeline for all variables',
           'container': flavour, 'variables': 2, 'slices_per_call': n,
           'ms_per_slice_previous_result_still_held': dt_retained / n * 1e3}
    del ds, out, a, b
    return res


def measure_configs(args, torch, device, m3, csr3, info3, ring, peak):
    """Every other BASELINE config, device-resident, CUDA events, median of 20 after 3 warm-ups.
    Inputs are either larger than L2 (ring of distinct slices) or L2 is flushed between
    iterations.  Returns ``{name: {ms, algorithmic_GBps, frac, kernel, ...}}``."""
    from pyremap_b200 import _cabi, synthetic as syn
    st = torch.cuda.current_stream(device).cuda_stream
    res = {}
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)

    def entry(name, ms, best, nbytes, kern, **extra):
        gbs = nbytes / (ms * 1e-3) / 1e9
        res[name] = dict({'ms': ms, 'best_ms': best, 'algorithmic_bytes': int(nbytes),
                          'algorithmic_GBps': gbs, 'frac': gbs / peak, 'kernel': kern}, **extra)

    def spmm(csr, x, y, K, nb, mode, n_src, n_dst, y_f32=False, kernel=_cabi.KERNEL_AUTO):
        code = _cabi.F64 if x.dtype == torch.float64 else _cabi.F32
        csr.spmm(x.data_ptr(), code, K, K, nb, n_src * K, y.data_ptr(), K, n_dst * K, mode,
                 THRESHOLD, stream=st, y_f32=y_f32, kernel=kernel)

    K = N_LEVELS
    n_a, n_b = m3.n_a, m3.n_b
    # ---- C3 variants on the resident ring (18.9 GB: larger than L2)
    y8 = torch.empty((BATCH, n_b, K), dtype=torch.float64, device=device)
    ms, best = time_launches(torch, lambda i: spmm(csr3, ring[i % RING], y8, K, 1, 2, n_a, n_b))
    entry('C3 masked, one slice per launch', ms, best, launch_bytes(info3, K, 1),
          kernel_name(csr3, mode=2), slices_per_launch=1)
    ring32 = ring.to(torch.float32)
    y32 = torch.empty((BATCH, n_b, K), dtype=torch.float32, device=device)
    ms, best = time_launches(torch, lambda i: spmm(csr3, ring32, y8, K, BATCH, 2, n_a, n_b))
    entry('C3 masked, float32 fields in, float64 out, x8', ms, best,
          launch_bytes(info3, K, BATCH, w_in=4), kernel_name(csr3, 'float', mode=2),
          slices_per_launch=BATCH)
    ms, best = time_launches(torch, lambda i: spmm(csr3, ring32, y32, K, BATCH, 2, n_a, n_b, y_f32=True))
    entry('C3 masked, float32 in and out, x8', ms, best,
          launch_bytes(info3, K, BATCH, w_in=4, w_out=4), kernel_name(csr3, 'float', mode=2),
          slices_per_launch=BATCH, tolerance='every element = float32(reference float64 result)')
    ms, best = time_launches(torch, lambda i: spmm(csr3, ring32, y32, K, BATCH, 2, n_a, n_b, y_f32=True,
                                                   kernel=_cabi.KERNEL_WROW_F32))
    entry('C3 masked, float32 in and out, float32 arithmetic (opt-in), x8', ms, best,
          launch_bytes(info3, K, BATCH, w_in=4, w_out=4), 'wrow_kernel<float,VEC=4,MODE=2,F32C>',
          slices_per_launch=BATCH,
          tolerance='NaN placement bit-exact (float64 denominator), values within 1e-6 relative')
    del ring32, y32
    if args.mode == 'masked':
        ring.nan_to_num_(nan=1.5)                   # the unmasked (frac_b) branch needs NaN-free data
    ms, best = time_launches(torch, lambda i: spmm(csr3, ring, y8, K, BATCH, 1, n_a, n_b))
    entry('C3 unmasked (frac_b branch), x8', ms, best, launch_bytes(info3, K, BATCH, with_fracb=True),
          kernel_name(csr3, mode=1), slices_per_launch=BATCH)
    del y8

    # ---- C1: 2 deg -> 1 deg bilinear, 10 levels, unmasked: launch-latency bound
    m1 = syn.make_c1()
    _, csr1, info1 = csr_of(m1, device.index)
    x1 = torch.randn((1, m1.n_a, 10), dtype=torch.float64, device=device)
    y1 = torch.empty((1, m1.n_b, 10), dtype=torch.float64, device=device)
    ms, best = time_launches(torch, lambda i: spmm(csr1, x1, y1, 10, 1, 1, m1.n_a, m1.n_b), reps=50,
                             flush=flush)
    entry('C1 2deg->1deg bilinear, K=10, frac_b branch', ms, best,
          launch_bytes(info1, 10, 1, with_fracb=True), kernel_name(csr1, vec=2, mode=1, K=10),
          latency_us=ms * 1e3, note='launch-latency bound (10 MB of traffic): the latency is the figure',
          l2='flushed between iterations')
    del csr1, x1, y1

    # ---- C2: MPAS-like 235k cells -> 0.5 deg, 60 levels x 12 months, land-masked
    m2 = syn.make_c2()
    _, csr2, info2 = csr_of(m2, device.index)
    lv = torch.from_numpy(syn.bathymetry_levels(m2.n_a, 60, seed=5)).to(device)
    nat = torch.empty((12, m2.n_a, 60), dtype=torch.float64, device=device).uniform_(-2.0, 30.0)
    nat.masked_fill_((torch.arange(60, device=device)[None, :] >= lv[:, None])[None], float('nan'))
    y2 = torch.empty((12, m2.n_b, 60), dtype=torch.float64, device=device)
    ms, best = time_launches(torch, lambda i: spmm(csr2, nat, y2, 60, 12, 2, m2.n_a, m2.n_b),
                             flush=flush)
    entry('C2 native (Time=12, nCells, nVertLevels=60) masked', ms, best,
          launch_bytes(info2, 60, 12), kernel_name(csr2, mode=2, K=60), slices_per_launch=12,
          l2='flushed between iterations')
    flat = nat.permute(1, 0, 2).reshape(1, m2.n_a, 720).contiguous()
    y2f = y2.view(1, m2.n_b, 720)
    ms, best = time_launches(torch, lambda i: spmm(csr2, flat, y2f, 720, 1, 2, m2.n_a, m2.n_b),
                             flush=flush)
    entry('C2 flat [nCells, K=720] masked', ms, best, launch_bytes(info2, 720, 1),
          kernel_name(csr2, mode=2, K=720), slices_per_launch=1, l2='flushed between iterations')
    del csr2, nat, flat, y2, y2f

    # ---- C4: 1 km -> 10 km stereographic, 121 entries per row, K = 1 and K = 4
    m4 = syn.make_c4()
    _, csr4, info4 = csr_of(m4, device.index)
    for K4 in (1, 4):
        x4 = torch.empty((2, m4.n_a, K4), dtype=torch.float64, device=device).uniform_(-2.0, 30.0)
        x4[:, ::97, :] = float('nan')
        y4 = torch.empty((1, m4.n_b, K4), dtype=torch.float64, device=device)
        ms, best = time_launches(torch, lambda i: spmm(csr4, x4[i % 2], y4, K4, 1, 2, m4.n_a, m4.n_b),
                                 flush=flush)
        entry(f'C4 1km->10km conservative (121 entries/row), K={K4}, masked', ms, best,
              launch_bytes(info4, K4, 1), kernel_name(csr4, vec=4 if K4 == 4 else 1, mode=2, K=K4),
              slices_per_launch=1, l2='flushed between iterations')
        del x4, y4
    del csr4, flush
    torch.cuda.empty_cache()
    return res


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

"""Deterministic synthetic mapping files and fields for the BASELINE.json configs.

Mapping-file *generation* (ESMF_RegridWeightGen / mbtempest, reference
``pyremap/remapper/build_map.py:8-212``) is out of scope and its binaries do not
exist on the GPU box, so the benchmark and the parity tests use analytic stand-ins
with the same sizes, sparsity structure and file layout (``S``, ``row``, ``col``
1-based, ``frac_b``, ``src_grid_dims``/``dst_grid_dims`` in Fortran order) as the
maps those tools write.  Everything is seeded; ``scale`` shrinks a config for
tests without changing its character.

C1  2 deg -> 1 deg lat-lon bilinear                (``make_c1``)
C2  MPAS-like 235k cells -> 0.5 deg conservative   (``make_c2``)
C3  MPAS-like 3.7M cells -> 10 km Antarctic stereo (``make_c3``)   [C5 = C3 x 365 slices]
C4  stereo 1 km -> 10 km conservative, 121 nnz/row (``make_c4``)
"""

from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

EARTH_RADIUS_KM = 6371.0


class SimpleDescriptor:
    """Duck-typed stand-in for ``pyremap.descriptor.MeshDescriptor``: the hot
    path reads only these four attributes (reference ``remap_numpy.py:67,113-132,198``)."""

    def __init__(self, dims, dim_sizes, coords=None, mesh_name='mesh'):
        self.dims = list(dims)
        self.dim_sizes = [int(s) for s in dim_sizes]
        self.coords = dict(coords or {})
        self.mesh_name = mesh_name
        self.regional = False


@dataclass
class SyntheticMap:
    name: str
    n_a: int
    n_b: int
    S: np.ndarray          # float64 [n_s]
    row: np.ndarray        # int32 [n_s], 1-based destination index
    col: np.ndarray        # int32 [n_s], 1-based source index
    frac_b: np.ndarray     # float64 [n_b]
    src_grid_dims: np.ndarray   # int32, Fortran order (fastest first)
    dst_grid_dims: np.ndarray
    src_descriptor: SimpleDescriptor
    dst_descriptor: SimpleDescriptor
    meta: dict = field(default_factory=dict)

    @property
    def n_s(self):
        return int(self.S.size)

    def save_npz(self, filename):
        np.savez(filename, S=self.S, row=self.row, col=self.col,
                 frac_b=self.frac_b, src_grid_dims=self.src_grid_dims,
                 dst_grid_dims=self.dst_grid_dims,
                 frac_a=np.zeros(self.n_a, dtype=np.float32))
        return filename


# --------------------------------------------------------------------------
# helpers
# --------------------------------------------------------------------------
def _latlon_coords(lat, lon, lat_name='lat', lon_name='lon'):
    return {lat_name: {'dims': (lat_name,), 'data': lat,
                       'attrs': {'units': 'degrees_north'}},
            lon_name: {'dims': (lon_name,), 'data': lon,
                       'attrs': {'units': 'degrees_east'}}}


def _unit_vectors(lat_deg, lon_deg):
    lat = np.deg2rad(lat_deg)
    lon = np.deg2rad(lon_deg)
    c = np.cos(lat)
    return np.stack([c * np.cos(lon), c * np.sin(lon), np.sin(lat)], axis=-1)


def _fibonacci_sphere(n):
    i = np.arange(n, dtype=np.float64) + 0.5
    z = -1.0 + 2.0 * i / n                      # south -> north
    golden = np.pi * (3.0 - np.sqrt(5.0))
    lon = np.rad2deg((i * golden) % (2 * np.pi)) - 180.0
    lat = np.rad2deg(np.arcsin(z))
    return lat, lon


def _ragged_neighbour_weights(dist, idx, radius, min_keep, rng):
    """Turn a k-nearest query into ragged positive weights.

    Keeps the neighbours closer than ``radius`` but never fewer than
    ``min_keep``; weight = compact bump of the distance times a little seeded
    jitter (so weights are not symmetric or exactly representable)."""
    k = dist.shape[1]
    rank = np.arange(k)[None, :]
    keep = (dist < radius) | (rank < min_keep)
    keep &= np.isfinite(dist)
    bump = np.clip(1.0 - (dist / (1.25 * radius)) ** 2, 0.05, None)
    w = np.where(keep, bump * rng.uniform(0.8, 1.2, size=dist.shape), 0.0)
    return keep, w


def _assemble(rows_keep, w, idx, frac_row, sort_cols=True):
    """Ragged [n_row, k] -> 1-based COO triplets with row sums == frac_row."""
    tot = w.sum(axis=1)
    scale = np.divide(frac_row, tot, out=np.zeros_like(tot), where=tot > 0)
    w = w * scale[:, None]
    live = rows_keep & (frac_row[:, None] > 0)
    r, j = np.nonzero(live)
    c = idx[r, j]
    s = w[r, j]
    if sort_cols:
        order = np.lexsort((c, r))
        r, c, s = r[order], c[order], s[order]
    return (s.astype(np.float64), (r + 1).astype(np.int32),
            (c + 1).astype(np.int32))


def _pseudo_land(lat_deg, lon_deg, seed, fraction):
    """Smooth pseudo-continents covering about ``fraction`` of the cells.
    Returns a 'land-ness' in [0,1]: 1 = fully land, 0 = open ocean."""
    rng = np.random.default_rng(seed)
    lat = np.deg2rad(lat_deg)
    lon = np.deg2rad(lon_deg)
    f = np.zeros(np.broadcast(lat, lon).shape)
    for _ in range(6):
        a, b = rng.integers(1, 5, size=2)
        ph1, ph2 = rng.uniform(0, 2 * np.pi, size=2)
        f = f + rng.uniform(0.5, 1.0) * np.sin(a * lon + ph1) * np.cos(b * lat + ph2)
    cut = np.quantile(f, 1.0 - fraction)
    width = 0.05 * (f.max() - f.min())
    return np.clip((f - cut) / width + 0.5, 0.0, 1.0)


# --------------------------------------------------------------------------
# C1
# --------------------------------------------------------------------------
def make_c1(src_res=2.0, dst_res=1.0, shuffle_triplets=False, seed=1):
    """Analytic bilinear lat-lon -> lat-lon map (4 entries per row, periodic in
    longitude, clamped at the poles); ``frac_b = 1``."""
    nlat_a, nlon_a = int(round(180 / src_res)), int(round(360 / src_res))
    nlat_b, nlon_b = int(round(180 / dst_res)), int(round(360 / dst_res))
    lat_a = -90 + src_res * (np.arange(nlat_a) + 0.5)
    lon_a = -180 + src_res * (np.arange(nlon_a) + 0.5)
    lat_b = -90 + dst_res * (np.arange(nlat_b) + 0.5)
    lon_b = -180 + dst_res * (np.arange(nlon_b) + 0.5)

    fi = (lat_b - lat_a[0]) / src_res
    i0 = np.clip(np.floor(fi).astype(np.int64), 0, nlat_a - 2)
    t = np.clip(fi - i0, 0.0, 1.0)
    fj = (lon_b - lon_a[0]) / src_res
    j0f = np.floor(fj)
    u = fj - j0f
    j0 = j0f.astype(np.int64) % nlon_a
    j1 = (j0 + 1) % nlon_a

    I0, J0 = np.meshgrid(i0, j0, indexing='ij')
    _, J1 = np.meshgrid(i0, j1, indexing='ij')
    T, U = np.meshgrid(t, u, indexing='ij')
    cols = np.stack([I0 * nlon_a + J0, I0 * nlon_a + J1,
                     (I0 + 1) * nlon_a + J0, (I0 + 1) * nlon_a + J1], axis=-1)
    wts = np.stack([(1 - T) * (1 - U), (1 - T) * U, T * (1 - U), T * U], axis=-1)
    n_b = nlat_b * nlon_b
    cols = cols.reshape(n_b, 4)
    wts = wts.reshape(n_b, 4)
    order = np.argsort(cols, axis=1, kind='stable')
    cols = np.take_along_axis(cols, order, axis=1)
    wts = np.take_along_axis(wts, order, axis=1)
    row = np.repeat(np.arange(n_b), 4)
    S, col = wts.ravel(), cols.ravel()
    if shuffle_triplets:
        p = np.random.default_rng(seed).permutation(S.size)
        S, row, col = S[p], row[p], col[p]
    src = SimpleDescriptor(['lat', 'lon'], [nlat_a, nlon_a],
                           _latlon_coords(lat_a, lon_a),
                           f'{src_res}x{src_res}degree')
    dst = SimpleDescriptor(['lat', 'lon'], [nlat_b, nlon_b],
                           _latlon_coords(lat_b, lon_b),
                           f'{dst_res}x{dst_res}degree')
    return SyntheticMap('C1', nlat_a * nlon_a, n_b, S.astype(np.float64),
                        (row + 1).astype(np.int32), (col + 1).astype(np.int32),
                        np.ones(n_b), np.array([nlon_a, nlat_a], np.int32),
                        np.array([nlon_b, nlat_b], np.int32), src, dst)


# --------------------------------------------------------------------------
# C2
# --------------------------------------------------------------------------
def make_c2(scale=1.0, seed=2, order='spiral', land_fraction=0.30):
    """MPAS-like quasi-uniform cells -> 0.5 deg lat-lon, conservative-like ragged
    rows (3..12 entries, mean ~7), ~30 % empty land rows with ``frac_b = 0`` and
    fractional ``frac_b`` along the coast."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    n_a = max(64, int(round(235160 * scale)))
    res = 0.5 / np.sqrt(scale)
    nlat_b = max(4, int(round(180 / res)))
    nlon_b = 2 * nlat_b
    res = 180.0 / nlat_b
    lat_a, lon_a = _fibonacci_sphere(n_a)
    if order == 'shuffle':
        p = rng.permutation(n_a)
        lat_a, lon_a = lat_a[p], lon_a[p]
    lat_b = -90 + res * (np.arange(nlat_b) + 0.5)
    lon_b = -180 + res * (np.arange(nlon_b) + 0.5)
    LAT, LON = np.meshgrid(lat_b, lon_b, indexing='ij')
    n_b = nlat_b * nlon_b

    tree = cKDTree(_unit_vectors(lat_a, lon_a))
    spacing = np.sqrt(4 * np.pi / n_a)
    radius = np.sqrt(7.0 / np.pi) * spacing
    dist, idx = tree.query(_unit_vectors(LAT.ravel(), LON.ravel()), k=12,
                           workers=-1)
    keep, w = _ragged_neighbour_weights(dist, idx, radius, 3, rng)
    land = _pseudo_land(LAT.ravel(), LON.ravel(), seed + 100, land_fraction)
    frac_b = np.where(land >= 1.0, 0.0, 1.0 - land)
    S, row, col = _assemble(keep, w, idx, frac_b)
    src = SimpleDescriptor(['nCells'], [n_a], {
        'lat_cell': {'dims': ('nCells',), 'data': lat_a, 'attrs': {}},
        'lon_cell': {'dims': ('nCells',), 'data': lon_a, 'attrs': {}}},
        f'synthMPAS{n_a}')
    dst = SimpleDescriptor(['lat', 'lon'], [nlat_b, nlon_b],
                           _latlon_coords(lat_b, lon_b), f'{res}x{res}degree')
    return SyntheticMap('C2', n_a, n_b, S, row, col, frac_b,
                        np.array([n_a], np.int32),
                        np.array([nlon_b, nlat_b], np.int32), src, dst,
                        {'lat_a': lat_a, 'lon_a': lon_a})


# --------------------------------------------------------------------------
# C3 (and C5)
# --------------------------------------------------------------------------
def _variable_resolution_rings(n_target, res_pole_km, res_equator_km):
    """Latitude rings whose spacing follows res(lat) (fine at the poles, coarse
    at the equator, like oRRS18to6); exactly ``n_target`` points, ordered ring by
    ring from the south pole northwards."""

    def rings(f):
        lats, counts = [], []
        lat = -90.0
        while True:
            c = np.cos(np.deg2rad(lat))
            res = f * (res_pole_km + (res_equator_km - res_pole_km) * c * c)
            lat_mid = lat + 0.5 * np.rad2deg(res / EARTH_RADIUS_KM)
            if lat_mid >= 90.0:
                break
            circ = 2 * np.pi * EARTH_RADIUS_KM * np.cos(np.deg2rad(lat_mid))
            lats.append(lat_mid)
            counts.append(max(1, int(round(circ / res))))
            lat = lat + np.rad2deg(res / EARTH_RADIUS_KM)
        return np.array(lats), np.array(counts)

    lo, hi = 0.05, 50.0
    for _ in range(60):
        mid = np.sqrt(lo * hi)
        if rings(mid)[1].sum() >= n_target:
            lo = mid
        else:
            hi = mid
    lats, counts = rings(lo)
    surplus = int(counts.sum() - n_target)
    # shave the surplus off the longest (equatorial) rings
    while surplus > 0:
        j = int(np.argmax(counts))
        take = min(surplus, max(1, counts[j] // 50))
        counts[j] -= take
        surplus -= take
    lat = np.repeat(lats, counts)
    start = np.concatenate([[0], np.cumsum(counts)[:-1]])
    within = np.arange(counts.sum()) - np.repeat(start, counts)
    phase = np.repeat((np.arange(counts.size) * 0.618034) % 1.0, counts)
    lon = ((within + phase) / np.repeat(counts, counts)) * 360.0 - 180.0
    return lat, lon


def _antarctic_stereo_xy(lat_deg, lon_deg, lat_ts=-71.0):
    """Spherical south-polar stereographic projection (km), true scale at lat_ts."""
    k = 1.0 + np.sin(np.deg2rad(abs(lat_ts)))
    rho = EARTH_RADIUS_KM * k * np.tan(np.pi / 4 + np.deg2rad(lat_deg) / 2)
    lon = np.deg2rad(lon_deg)
    return rho * np.sin(lon), rho * np.cos(lon)


def make_c3(scale=1.0, seed=3, order='rings'):
    """MPAS-like variable-resolution ocean mesh (3 693 225 cells at scale 1) ->
    Antarctic stereographic grid, 6000 x 5000 km at 10 km (601 x 501); 3..10
    entries per ocean row, rows over the continent empty (``frac_b = 0``).
    ``order``: numbering of the source cells -- 'rings' (latitude rings from the south pole),
    'blocks' (compact blocks in random order, graph-partition-like) or 'shuffle' (random)."""
    from scipy.spatial import cKDTree
    rng = np.random.default_rng(seed)
    n_a = max(2000, int(round(3693225 * scale)))
    lin = 1.0 / np.sqrt(scale)
    lat_a, lon_a = _variable_resolution_rings(n_a, 6.0 * lin, 18.0 * lin)
    if order == 'shuffle':
        p = rng.permutation(n_a)
        lat_a, lon_a = lat_a[p], lon_a[p]
    elif order == 'blocks':
        # graph-partition-like numbering (how MPAS meshes come out of METIS-style tools): cells
        # of one compact block are contiguous, the blocks themselves in no geographic order
        nb_lat, nb_lon = 48, 96
        bi = np.minimum(((lat_a + 90.0) / 180.0 * nb_lat).astype(np.int64), nb_lat - 1)
        bj = np.minimum(((lon_a + 180.0) / 360.0 * nb_lon).astype(np.int64), nb_lon - 1)
        block_rank = rng.permutation(nb_lat * nb_lon)[bi * nb_lon + bj]
        p = np.argsort(block_rank, kind='stable')
        lat_a, lon_a = lat_a[p], lon_a[p]
    dx = 10.0 * lin
    nx = int(6000.0 / dx) + 1
    ny = int(5000.0 / dx) + 1
    x = dx * (np.arange(nx) - (nx - 1) / 2)
    y = dx * (np.arange(ny) - (ny - 1) / 2)
    X, Y = np.meshgrid(x, y, indexing='xy')      # [ny, nx]
    n_b = nx * ny

    near = np.nonzero(lat_a < -40.0)[0]
    xs, ys = _antarctic_stereo_xy(lat_a[near], lon_a[near])
    tree = cKDTree(np.stack([xs, ys], axis=-1))
    dist, idx = tree.query(np.stack([X.ravel(), Y.ravel()], axis=-1), k=10,
                           workers=-1)
    idx = near[np.minimum(idx, near.size - 1)]
    keep, w = _ragged_neighbour_weights(dist, idx, 1.0 * dx, 3, rng)

    theta = np.arctan2(Y.ravel(), X.ravel())
    r = np.hypot(X.ravel(), Y.ravel())
    coast = (1900.0 + 250.0 * np.sin(3 * theta + 0.7)
             + 180.0 * np.cos(5 * theta - 1.1) - 500.0 * (np.cos(theta + 2.4) > 0.8))
    landness = np.clip((coast - r) / (2.0 * dx) + 0.5, 0.0, 1.0)
    frac_b = np.where(landness >= 1.0, 0.0, 1.0 - landness)
    S, row, col = _assemble(keep, w, idx, frac_b)
    src = SimpleDescriptor(['nCells'], [n_a], {}, f'synth_oRRS18to6_{n_a}')
    dst = SimpleDescriptor(['y', 'x'], [ny, nx], {
        'x': {'dims': ('x',), 'data': x * 1e3, 'attrs': {'units': 'm'}},
        'y': {'dims': ('y',), 'data': y * 1e3, 'attrs': {'units': 'm'}}},
        f'{dx:g}km_Antarctic_stereo')
    return SyntheticMap('C3', n_a, n_b, S, row, col, frac_b,
                        np.array([n_a], np.int32), np.array([nx, ny], np.int32),
                        src, dst, {'lat_a': lat_a})


# --------------------------------------------------------------------------
# C4
# --------------------------------------------------------------------------
def make_c4(scale=1.0, ratio=10):
    """Exact conservative overlap weights of two aligned Cartesian grids
    (1 km -> 10 km at scale 1): a destination cell overlaps ``ratio + 1`` source
    cells per direction, the two outermost with half weight -> 121 entries per
    interior row; ``frac_b = 1``."""
    nx_b = max(4, int(round(600 * scale))) + 1
    ny_b = max(4, int(round(500 * scale))) + 1
    nx_a = (nx_b - 1) * ratio + 1
    ny_a = (ny_b - 1) * ratio + 1
    half = ratio // 2
    off = np.arange(-half, half + 1)
    w1 = np.ones(off.size)
    if ratio % 2 == 0:
        w1[0] = w1[-1] = 0.5

    def axis_tables(n_b_axis, n_a_axis):
        pos = (np.arange(n_b_axis) * ratio)[:, None] + off[None, :]
        ok = (pos >= 0) & (pos < n_a_axis)
        w = np.where(ok, w1[None, :], 0.0)
        w = w / w.sum(axis=1, keepdims=True)
        return pos, ok, w

    py, oky, wy = axis_tables(ny_b, ny_a)
    px, okx, wx = axis_tables(nx_b, nx_a)
    m = off.size
    # [ny_b, nx_b, m(y), m(x)]
    ok = oky[:, None, :, None] & okx[None, :, None, :]
    cols = (py[:, None, :, None].astype(np.int64) * nx_a
            + px[None, :, None, :].astype(np.int64))
    wts = wy[:, None, :, None] * wx[None, :, None, :]
    n_b = ny_b * nx_b
    ok = ok.reshape(n_b, m * m)
    r, j = np.nonzero(ok)
    col = np.broadcast_to(cols, (ny_b, nx_b, m, m)).reshape(n_b, m * m)[r, j]
    S = np.broadcast_to(wts, (ny_b, nx_b, m, m)).reshape(n_b, m * m)[r, j]
    x_b = np.arange(nx_b) * float(ratio)
    y_b = np.arange(ny_b) * float(ratio)
    src = SimpleDescriptor(['y', 'x'], [ny_a, nx_a], {}, '1km_stereo')
    dst = SimpleDescriptor(['y', 'x'], [ny_b, nx_b], {
        'x': {'dims': ('x',), 'data': x_b, 'attrs': {}},
        'y': {'dims': ('y',), 'data': y_b, 'attrs': {}}}, '10km_stereo')
    return SyntheticMap('C4', ny_a * nx_a, n_b, S.astype(np.float64),
                        (r + 1).astype(np.int32), (col + 1).astype(np.int32),
                        np.ones(n_b), np.array([nx_a, ny_a], np.int32),
                        np.array([nx_b, ny_b], np.int32), src, dst)


# --------------------------------------------------------------------------
# fields
# --------------------------------------------------------------------------
def bathymetry_levels(n_cells, n_levels, seed):
    """maxLevelCell-like clipped log-normal: cell ``c`` is valid for levels
    ``< max_level[c]`` (0 = a dry cell)."""
    rng = np.random.default_rng(seed)
    lv = np.exp(rng.normal(np.log(0.55 * n_levels), 0.6, size=n_cells))
    lv = np.clip(np.round(lv), 0, n_levels).astype(np.int32)
    lv[rng.random(n_cells) < 0.02] = 0
    return lv


def ocean_field(n_cells, n_levels, seed, dtype=np.float64, max_level=None):
    """Temperature-like values in [-2, 30], NaN below ``max_level`` if given.
    Shape ``[n_cells, n_levels]`` (levels fastest, MPAS native)."""
    rng = np.random.default_rng(seed)
    f = rng.uniform(-2.0, 30.0, size=(n_cells, n_levels)).astype(dtype)
    if max_level is not None:
        f[np.arange(n_levels)[None, :] >= max_level[:, None]] = np.nan
    return f

"""``Remapper``: the drop-in boundary for pyremap's weight-application path.

Constructor signature, attribute names and ``remap_numpy`` follow
``/root/reference/pyremap/remapper/remapper.py:74-137,508-532`` (pyremap 2.4.0).
Everything about *making* a mapping file (``src_from_*``/``dst_from_*``,
``build_map``) and the file-to-file NCO path (``ncremap``) stays with pyremap on
the CPU and is out of scope here; those methods raise ``NotImplementedError``
that says so instead of silently doing something else.
"""

from __future__ import annotations

from .remap_numpy import _remap_numpy, remap_array


class Remapper:
    """Apply an existing SCRIP/ESMF mapping file to fields on an NVIDIA B200.

    Attributes (same meaning as in pyremap): ``ntasks``, ``map_filename``,
    ``method``, ``src_descriptor``, ``dst_descriptor``, ``map_tool``,
    ``parallel_exec``, ``use_tmp``, ... and the lazily filled ``_ds_map`` /
    ``_matrix`` cache.  Added: ``device`` (CUDA device index or ``None`` for the
    current one).
    """

    def __init__(
        self,
        ntasks=1,
        map_filename=None,
        method='bilinear',
        src_descriptor=None,
        dst_descriptor=None,
        map_tool='esmf',
        parallel_exec='mpirun',
        use_tmp=True,
    ):
        self.ntasks = ntasks
        self.src_grid_info = dict()
        self.dst_grid_info = dict()
        self.map_filename = map_filename
        self.method = method
        self.use_tmp = use_tmp
        self.expand_dist = None
        self.expand_factor = None
        self.src_scrip_filename = 'src_mesh.nc'
        self.dst_scrip_filename = 'dst_mesh.nc'
        self.format = 'NETCDF3_64BIT_DATA'
        self.src_descriptor = src_descriptor
        self.dst_descriptor = dst_descriptor
        self.map_tool = map_tool
        self.esmf_path = None
        self.moab_path = None
        self.parallel_exec = parallel_exec
        self._ds_map = None
        self._matrix = None
        self.device = None

    # ---- the hot path -------------------------------------------------
    def remap_numpy(self, ds, renormalization_threshold=None):
        """Remap an ``xarray.Dataset``/``DataArray`` in memory
        (reference remapper.py:508-532).

        ``renormalization_threshold``: minimum weight of a destination cell
        after remapping below which it is masked out, or ``None`` for no
        renormalization and masking.  Returns the same type as ``ds``.
        """
        return _remap_numpy(self, ds, renormalization_threshold)

    #: pyremap 1.x spelling of the same call
    remap = remap_numpy

    def remap_array(self, field, remap_axes, renormalization_threshold=None,
                    return_torch=False, out_dtype=None, out=None, mode='auto', arithmetic=None):
        """Array-level entry: numpy array or CUDA tensor in, NaN-filled float64
        out (``return_torch=True`` keeps the result on the device;
        ``out_dtype=np.float32`` rounds the float64 result to float32 on the GPU;
        ``out=`` a preallocated host result, ideally pinned, for host inputs;
        ``mode='masked'|'fracb'`` imposes the branch the reference would pick for the
        whole variable when ``field`` is only a part of it; ``arithmetic='float32'``: float32
        products and sums for float32 fields with float32 results, within 1e-6 of the
        reference, NaN placement exact)."""
        return remap_array(self, field, remap_axes, renormalization_threshold,
                           return_torch=return_torch, out_dtype=out_dtype, out=out, mode=mode,
                           arithmetic=arithmetic)

    # ---- out of scope (CPU, stays with pyremap) -----------------------
    def build_map(self, logger=None):
        raise NotImplementedError(
            'mapping-file generation (ESMF_RegridWeightGen / mbtempest) stays '
            'on the CPU with pyremap; pyremap_b200 only applies existing maps')

    def remap_file(self, in_filename, out_filename, variable_list=None, overwrite=False,
                   renormalize=None, logger=None, replace_mpas_fill=False, parallel_exec=None):
        """File-to-file remap on the GPU with the arguments of the reference's
        ``Remapper.ncremap`` (remapper.py:434-506); see :mod:`pyremap_b200.remap_file`."""
        from .remap_file import remap_file
        return remap_file(self, in_filename, out_filename, variable_list=variable_list,
                          overwrite=overwrite, renormalize=renormalize, logger=logger,
                          replace_mpas_fill=replace_mpas_fill, parallel_exec=parallel_exec)

    #: the reference's name of the file-to-file call; here it runs on the GPU instead of
    #: launching NCO (same arguments, ``parallel_exec`` ignored)
    ncremap = remap_file

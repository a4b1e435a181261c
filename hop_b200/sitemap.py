"""Grid / Density with the site map: the stage-2 drop-in for hop.sitemap.

Public surface follows hop/sitemap.py (Grid 53-357, Density 361-1631): attributes ``grid``,
``edges``, ``P``, ``unit``, ``metadata``, ``map`` (int16), ``sites``, ``site_properties``, derived
``delta``/``midpoints``/``origin``; methods ``make_density``, ``convert_density``, ``map_sites``,
``site_insert_bulk``, ``site_remove_bulk``, ``site_insert_nobulk``, ``has_bulk``, ``site_labels``,
``site_volume``, ``site_occupancy``, ``save``/``load``, ``export`` (OpenDX).

Underneath, the density lives in HBM as float64 and the site map as int32 labels; thresholding,
connected-component labelling, size ranking, bulk insertion and the per-site reductions are CUDA
kernels (csrc/label.cu).  Host numpy views are materialised lazily when an attribute is read.
There is no CPU implementation of those steps in this package.
"""
from __future__ import annotations

import warnings

import numpy

from . import constants
from .constants import SITELABEL
from .exceptions import InconsistentDataWarning, MissingDataWarning
from .utilities import DefaultDict, Saveable, asiterable, flatten


def _torch():
    import torch

    return torch


class SiteCells:
    """One site: a read-only sequence of (i, j, k) cell index tuples backed by an (n, 3) array."""

    __slots__ = ("_ijk",)

    def __init__(self, ijk):
        self._ijk = numpy.asarray(ijk, dtype=numpy.int64).reshape(-1, 3)

    def __len__(self):
        return self._ijk.shape[0]

    def __iter__(self):
        return iter(map(tuple, self._ijk.tolist()))

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [tuple(r) for r in self._ijk[i].tolist()]
        return tuple(self._ijk[i].tolist())

    def __eq__(self, other):
        try:
            return list(self) == [tuple(x) for x in other]
        except TypeError:
            return NotImplemented

    def __repr__(self):
        return "SiteCells(%d cells)" % len(self)

    def asarray(self):
        return self._ijk


class Grid(Saveable):
    """A multidimensional histogram or density on a regular grid (hop/sitemap.py:53-357)."""

    _saved_attributes = ['grid', 'edges', 'P', 'unit', 'metadata']
    parameters_default = {'isDensity': False}
    unit_default = {'length': 'Angstrom', 'density': None}

    def __init__(self, grid=None, edges=None, filename=None, dxfile=None, parameters=None, unit=None, metadata=None):
        parameters = DefaultDict(self.parameters_default, parameters)
        unit = DefaultDict(self.unit_default, unit)
        metadata = DefaultDict({}, metadata)
        self._dxfile = dxfile
        self.P = parameters
        self.metadata = metadata
        self.unit = unit
        self._check_set_unit(unit)
        self._grid_host = None
        self._grid_dev = None
        self._tables = None
        if not (grid is None or edges is None):
            self.grid = grid
            self.edges = edges
        elif filename:
            super(Grid, self).__init__(filename=filename)
            self._update()
            return
        elif dxfile:
            self.importdx(dxfile)
            return
        else:
            raise ValueError("no input data---eg (grid,edges) or filename---for grid given")
        self._update()

    # ---- the grid lives on the host, on the device, or both ------------------------------------
    @property
    def grid(self):
        if self._grid_host is None and self._grid_dev is not None:
            self._grid_host = self._grid_dev.cpu().numpy()
        if self._grid_dev is not None:
            # a mirrored array is read-only: assign `obj.grid = new` to change the density
            self._grid_host.flags.writeable = False
        return self._grid_host

    @grid.setter
    def grid(self, value):
        t = _torch()
        if isinstance(value, t.Tensor):
            self._grid_dev = value if value.dtype == t.float64 else value.double()
            self._grid_host = None
        else:
            self._grid_host = numpy.array(value, dtype=numpy.float64)
            self._grid_dev = None

    def _device_grid(self):
        """float64 CUDA tensor of the grid (uploaded on first use)."""
        if self._grid_dev is None:
            t = _torch()
            if not t.cuda.is_available():
                raise RuntimeError("hop_b200 needs a CUDA device; there is no CPU path for this operation")
            self._grid_dev = t.from_numpy(numpy.ascontiguousarray(self._grid_host)).cuda()
        return self._grid_dev

    def _grid_tables(self):
        from .kernels import GridTables

        if self._tables is None:
            self._tables = GridTables(self.edges)
        return self._tables

    def _get_saved_attribute(self, attr):
        if attr == 'grid':
            g = self.grid
            return None if g is None else numpy.array(g)
        return self.__dict__[attr]

    def _set_loaded_attribute(self, attr, value):
        if attr == 'grid':
            self.grid = value
        else:
            self.__dict__[attr] = value

    def __getstate__(self):
        return super(Grid, self).__getstate__()

    def __setstate__(self, data):
        self._grid_host = None
        self._grid_dev = None
        self._tables = None
        for k, v in data.items():
            self._set_loaded_attribute(k, v)

    # ---- derived data ---------------------------------------------------------------------------
    def _update(self):
        """delta, midpoints, origin (centre of cell 0,0,0) -- hop/sitemap.py:153-165."""
        self.delta = numpy.diag([(e[-1] - e[0]) / (len(e) - 1) for e in self.edges])
        self.midpoints = [0.5 * (e[:-1] + e[1:]) for e in self.edges]
        self.origin = [m[0] for m in self.midpoints]
        self._tables = None
        if self.P['isDensity'] and not self.unit['density']:
            raise ValueError("For a density, unit['density'] must be set.")

    def _check_set_unit(self, u):
        try:
            for unit_type, value in u.items():
                if value is None:
                    self.unit[unit_type] = None
                    continue
                try:
                    constants.conversion_factor[unit_type][value]
                    self.unit[unit_type] = value
                except KeyError:
                    raise ValueError('Unit ' + str(value) + ' of type ' + str(unit_type) + ' is not recognized.')
        except AttributeError:
            raise ValueError('"unit" must be a dictionary with keys "length" and "density".')
        if 'length' not in self.unit:
            raise ValueError('"unit" must contain a unit for "length".')
        if 'density' not in self.unit:
            self.unit['density'] = None

    def make_density(self):
        """Histogram (counts per cell) -> density (counts per volume): divide by the bin widths along
        axis 0, 1, 2 in turn (hop/sitemap.py:196-216)."""
        if self.P['isDensity']:
            raise RuntimeError("Grid is already a density.")
        t = _torch()
        g = self._device_grid()
        D = len(self.edges)
        for i in range(D):
            shape = [1] * D
            w = numpy.diff(self.edges[i])
            shape[i] = len(w)
            g.div_(t.from_numpy(w.reshape(shape)).to(g.device))
        self._grid_host = None
        self.P['isDensity'] = True
        self.unit['density'] = self.unit['length'] + "^{-3}"

    def convert_length(self, unit='Angstrom'):
        if unit == self.unit['length']:
            return
        cvnfact = constants.get_conversion_factor('length', self.unit['length'], unit)
        self.edges = [x * cvnfact for x in self.edges]
        self.unit['length'] = unit
        self._update()

    def convert_density(self, unit='Angstrom^{-3}'):
        """Convert the density to the physical units given by unit (hop/sitemap.py:236-264)."""
        if not self.P['isDensity']:
            raise RuntimeError('The grid is not a density so convert_density() makes no sense.')
        if unit == self.unit['density']:
            return
        factor = constants.get_conversion_factor('density', self.unit['density'], unit)
        self._device_grid().mul_(factor)
        self._grid_host = None
        self.unit['density'] = unit

    def centers(self):
        """Coordinates of the centers of all grid cells as an iterator."""
        for idx in numpy.ndindex(self.grid.shape):
            yield numpy.sum(self.delta * numpy.asarray(idx), axis=0) + self.origin

    # ---- OpenDX -----------------------------------------------------------------------------------
    def export(self, filename=None, format="dx"):
        filename = self.filename(filename)
        if format == "dx":
            self._export_dx(filename)
        elif format == "python":
            self.save(filename)
        else:
            raise NotImplementedError("Exporting to format " + str(format) + " is not implemented.")

    def _export_dx(self, filename):
        """Regular-grid OpenDX file, data in C order (z fastest) -- hop/sitemap.py:311-345."""
        from .dx import write_dx

        comments = ['OpenDX density file written by hop_b200',
                    'Unit of   length    is ' + str(self.unit['length']) + '.',
                    'Unit of   density   is ' + str(self.unit['density']) + '-based.',
                    'Data is written in C array order: In grid[x,y,z] the axis z is fastest',
                    'varying, then y, then finally x, i.e. z is the innermost loop.']
        if self.metadata:
            comments.append('Meta data stored with the python Grid object:')
        for k in self.metadata:
            comments.append('   ' + str(k) + ' = ' + str(self.metadata[k]))
        write_dx(filename + '.dx', self.grid, self.origin, numpy.diag(self.delta), comments)

    def importdx(self, dxfile):
        from .dx import read_dx

        grid, edges = read_dx(dxfile)
        self.__init__(grid=grid, edges=edges, parameters=self.P, unit=self.unit, metadata=self.metadata, dxfile=dxfile)

    def __repr__(self):
        grid_type = 'density' if self.P['isDensity'] else 'histogram'
        try:
            fn = 'and default filename "' + str(self.filename()) + '"'
        except ValueError:
            fn = ''
        return '<hop_b200.sitemap.Grid ' + grid_type + ' with ' + str(tuple(self.grid.shape)) + ' bins ' + fn + '>'


class Density(Grid):
    """Density annotated with hydration sites (hop/sitemap.py:361-1631).

    A site is a set of at least ``MINsite`` cells with density >= threshold that are connected
    through the offsets {0,1}^3 (and their negatives).  Label 0 is the interstitial, 1 the largest
    site (or the bulk site once one is inserted), then sites in order of decreasing volume.
    """

    _saved_attributes = Grid._saved_attributes + ['map', 'sites', 'site_properties', 'equivalent_sites_index']
    parameters_default = {'isDensity': False, 'threshold': None, 'MINsite': 1}
    unit_default = {'length': 'Angstrom', 'density': None}

    def __init__(self, grid=None, edges=None, filename=None, dxfile=None, parameters=None, unit=None, metadata=None):
        parameters = DefaultDict(self.parameters_default, parameters)
        if parameters['MINsite'] < 1:
            raise ValueError('MINsite must be > 0 (site must contain at least one cell)')
        if parameters['MINsite'] > 2:
            raise NotImplementedError("Currently only MINsite == 1 or 2 is supported.")
        unit = DefaultDict(self.unit_default, unit)
        metadata = DefaultDict({}, metadata)
        self._reset_sites()
        super(Density, self).__init__(grid=grid, edges=edges, filename=filename, dxfile=dxfile,
                                      parameters=parameters, unit=unit, metadata=metadata)
        if not self.P['isDensity']:
            self.make_density()
        self.unit['threshold'] = self.unit['density']

    # ---- site state: device labels first, host structures on demand --------------------------------
    def _reset_sites(self):
        self._own = None          # int32 [nx,ny,nz]: label of the hydration site owning the cell, else 0
        self._member = None       # uint8: 1 where the cell belongs to the inserted bulk site
        self._map16 = None        # int16 painted map on the device
        self._n_labels = None     # len(sites)
        self._map_host = None
        self._sites_host = None
        self._props_host = None
        self._bulk_is_empty = False
        self.equivalent_sites_index = None

    @classmethod
    def _from_device_counts(cls, counts, edges, n_frames, parameters=None, unit=None, metadata=None):
        """Histogram counts (CUDA uint32) -> density in Angstrom^-3 without leaving the device
        (DensityCollector.finish + Grid.make_density, hop/density.py:133-137, hop/sitemap.py:196-216)."""
        from . import kernels as K

        self = cls.__new__(cls)
        parameters = DefaultDict(cls.parameters_default, parameters)
        if parameters['MINsite'] < 1:
            raise ValueError('MINsite must be > 0 (site must contain at least one cell)')
        if parameters['MINsite'] > 2:
            raise NotImplementedError("Currently only MINsite == 1 or 2 is supported.")
        self._reset_sites()
        self._dxfile = None
        self.P = parameters
        self.metadata = DefaultDict({}, metadata)
        self.unit = DefaultDict(cls.unit_default, unit)
        self._check_set_unit(self.unit)
        self._grid_host = None
        self._tables = None
        self.edges = [numpy.asarray(e, dtype=numpy.float64) for e in edges]
        self._update()
        self._grid_dev = K.density_from_counts(self._grid_tables(), counts, n_frames, 1.0)
        self.P['isDensity'] = True
        self.unit['density'] = self.unit['length'] + "^{-3}"
        self.unit['threshold'] = self.unit['density']
        return self

    @property
    def map(self):
        if self._map_host is None and self._map16 is not None:
            self._map_host = self._map16.cpu().numpy()
        return self._map_host

    @map.setter
    def map(self, value):
        self._map_host = None if value is None else numpy.asarray(value, dtype=numpy.int16)
        self._map16 = None

    def _device_map(self):
        """int16 CUDA tensor of the painted site map."""
        if self._map16 is None:
            if self._map_host is None:
                raise ValueError('The map has not been generated. Run map_sites() first.')
            self._map16 = _torch().from_numpy(numpy.ascontiguousarray(self._map_host)).cuda()
        return self._map16

    @property
    def sites(self):
        if self._sites_host is None and self._own is not None:
            from . import kernels as K

            cells, offsets = K.site_cells(self._own, self._member, self._n_labels)
            cells = cells.cpu().numpy()
            offsets = offsets.cpu().numpy()
            shape = tuple(self._own.shape)
            ijk = numpy.stack(numpy.unravel_index(cells, shape), axis=1) if len(cells) else numpy.zeros((0, 3), int)
            self._sites_host = [SiteCells(ijk[offsets[l]:offsets[l + 1]]) for l in range(self._n_labels)]
        return self._sites_host

    @sites.setter
    def sites(self, value):
        # manual manipulation of the site list: the device labels no longer describe it
        self._sites_host = value
        self._own = None
        self._member = None
        self._n_labels = None if value is None else len(value)
        self._props_host = None

    @property
    def site_properties(self):
        if self._props_host is None and (self._own is not None or self._sites_host is not None):
            self._annotate_sites()
        return self._props_host

    @site_properties.setter
    def site_properties(self, value):
        self._props_host = value

    def _get_saved_attribute(self, attr):
        if attr == 'map':
            return self.map
        if attr == 'sites':
            s = self.sites
            return None if s is None else [list(x) for x in s]
        if attr == 'site_properties':
            return self.site_properties
        if attr == 'equivalent_sites_index':
            return self.equivalent_sites_index
        return super(Density, self)._get_saved_attribute(attr)

    def _set_loaded_attribute(self, attr, value):
        if attr == 'map':
            self.map = value
        elif attr == 'sites':
            self._sites_host = None if value is None else [SiteCells(s) for s in value]
            self._n_labels = None if value is None else len(value)
        elif attr == 'site_properties':
            self._props_host = value
        elif attr == 'equivalent_sites_index':
            self.equivalent_sites_index = value
        else:
            super(Density, self)._set_loaded_attribute(attr, value)

    def __setstate__(self, data):
        self._reset_sites()
        super(Density, self).__setstate__(data)

    def load(self, filename=None, merge=False):
        self._reset_sites()
        return super(Density, self).load(filename, merge)

    # ---- stage 2 ------------------------------------------------------------------------------------
    def map_sites(self, threshold=None):
        """Find regions of connected density >= threshold and label them by decreasing size
        (hop/sitemap.py:480-521).  The threshold is in the units of the density."""
        from . import kernels as K

        self.unit['threshold'] = self.unit['density']
        if threshold:
            self.P['threshold'] = threshold
        else:
            try:
                threshold = self.P['threshold']
                if threshold is None:
                    raise ValueError
            except (KeyError, ValueError):
                raise ValueError("A threshold density is required to map the sites.")
        if 'bulk_site' in self.P:
            del self.P['bulk_site']
            self.P.pop('bulk_threshold', None)
            warnings.warn("Removed bulk site in order to map sites. Add it again when this is done!")
        self._reset_sites()
        labels, n_sites, _ = K.label_sites(self._device_grid(), float(self.P['threshold']), int(self.P['MINsite']))
        self._own = labels
        self._member = None
        self._n_labels = n_sites + 1
        self._map16 = K.labels_to_map16(labels)

    def _require_device_sites(self):
        """Device labels for this density's sites; rebuilt from a manually edited site list."""
        if self._own is not None:
            return
        if self._sites_host is None:
            raise ValueError('The map has not been generated. Run map_sites() first.')
        t = _torch()
        shape = tuple(self.grid.shape)
        own = numpy.zeros(shape, dtype=numpy.int32)
        for label, site in enumerate(self._sites_host):
            if len(site):
                a = numpy.asarray(site if not isinstance(site, SiteCells) else site.asarray())
                own[a[:, 0], a[:, 1], a[:, 2]] = label  # later sites overwrite (sitemap.py:1584-1587)
        self._own = t.from_numpy(own).cuda()
        self._member = None
        self._n_labels = len(self._sites_host)

    def site_insert_bulk(self, bulkdensity, bulklabel=SITELABEL['bulk'], force=False):
        """Insert site `bulklabel` of another density (same grid) as the bulk site, label 1; all
        sites of this density move up by one label (hop/sitemap.py:877-918)."""
        from . import kernels as K

        if 'bulk_site' in self.P and not force:
            raise ValueError('The density already contains a bulk site. Use force=True to override.')
        if self.map is None or bulkdensity.map is None:
            raise ValueError('Both densities must have had their map computed.')
        if self.map.shape != bulkdensity.map.shape:
            raise ValueError("The bulk density was defined on a different grid than this density.")
        if self.unit['threshold'] != bulkdensity.unit['threshold']:
            warnings.warn("The unit for the density (%s) is different from the unit for the bulk density (%s).\n"
                          % (self.unit['threshold'], bulkdensity.unit['threshold']), category=InconsistentDataWarning)
        self._require_device_sites()
        bulk_labels = None
        if isinstance(bulkdensity, Density):
            bulkdensity._require_device_sites()
            if bulkdensity._n_labels > bulklabel:
                bulk_labels = bulkdensity._own
            else:
                raise IndexError("bulk density has no site %d" % bulklabel)
        self._member, self._map16 = K.insert_bulk(self._own, bulk_labels, int(bulklabel))
        self._n_labels += 1
        self._map_host = None
        self._sites_host = None
        self._props_host = None
        self.P['bulk_site'] = SITELABEL['bulk']
        self.P['bulk_threshold'] = bulkdensity.P['threshold']

    def site_remove_bulk(self, force=False):
        """Cleanup bulk site."""
        self._site_remove_bulk(force=force)

    def _site_remove_bulk(self, force=False):
        if 'bulk_site' not in self.P and not force:
            raise ValueError("No bulk site is recorded but force=True would do it")
        self._require_device_sites()
        t = _torch()
        own = self._own
        own.sub_((own > 0).to(own.dtype))   # every hydration site moves back down by one label
        self._member = None
        self._n_labels -= 1
        from . import kernels as K

        self._map16 = K.labels_to_map16(own)
        self._map_host = None
        self._sites_host = None
        self._props_host = None
        self.P.pop('bulk_site', None)
        self.P.pop('bulk_threshold', None)

    def site_insert_nobulk(self):
        """Insert an empty bulk site for cases when this is convenient (hop/sitemap.py:933-943)."""
        class Nobulk:
            def __init__(self, dens):
                self.map = numpy.empty_like(dens.map)
                self.unit = dens.unit
                self.P = {'threshold': None}
                self.sites = {SITELABEL['bulk']: ()}
        self.site_insert_bulk(Nobulk(self))

    def has_bulk(self):
        """True if a bulk site has been inserted."""
        try:
            return self.P['bulk_site'] == SITELABEL['bulk']
        except KeyError:
            return False

    # ---- annotation -----------------------------------------------------------------------------------
    def _annotate_sites(self):
        """site_properties: label, volume, occupancy avg/std, centre per site
        (hop/sitemap.py:1589-1612, 601-627); the reductions run on the device."""
        from . import kernels as K

        self._require_device_sites()
        n = self._n_labels
        count, mean, std, center = K.site_properties(self._grid_tables(), self._device_grid(), self._own,
                                                     self._member, n)
        count = count.cpu().numpy()
        mean = mean.cpu().numpy()
        std = std.cpu().numpy()
        center = center.cpu().numpy()
        det = numpy.linalg.det(self.delta)
        factor = constants.get_conversion_factor('density', self.unit['density'], self.unit['length'] + "^{-3}")
        V = count * (factor * det)
        labels = numpy.arange(n)
        centers = numpy.empty(n, dtype=object)
        for l in range(n):
            centers[l] = center[l].copy() if count[l] > 0 else None
        names = numpy.empty(n, dtype=object)
        names[:] = ''
        self._props_host = numpy.rec.fromarrays(
            [labels, det * count.astype(numpy.float64), mean * V, std * V, centers,
             numpy.zeros(n, dtype=int), numpy.zeros(n, dtype=int), names],
            names='label,volume,occupancy_avg,occupancy_std,center,equivalence_label,equivalence_site,equivalence_name')

    def site_occupancy(self, **labelargs):
        """labels, <N>, std(N) of the selected sites (hop/sitemap.py:563-583)."""
        labels = self.site_labels(**labelargs)
        props = self.site_properties.take(labels)
        return labels, props.occupancy_avg, props.occupancy_std

    def site_volume(self, **labelargs):
        """labels, volumes of the selected sites (hop/sitemap.py:585-598)."""
        labels = self.site_labels(**labelargs)
        return labels, self.site_properties.take(labels).volume

    def site_labels(self, include='default', exclude='default'):
        """Site labels selected by keywords / integers (hop/sitemap.py:629-789).

        include: 'all' | 'default' (all but interstitial and subsites) | 'sites' (also without bulk)
        | 'subsites' | 'equivalencesites' | int | list of those; exclude: 'default' (=
        ['interstitial','subsites']) | None | 'interstitial' | 'bulk' | 'subsites' |
        'equivalencesites'.  Exclusions are applied after inclusions."""
        incl_kw = {None: ['default'], 'default': ['default'], 'all': ['all'], 'sites': ['sites'],
                   'subsites': ['subsites'], 'equivalencesites': ['equivalencesites']}
        excl_kw = {None: [], 'default': ['interstitial', 'subsites'], 'interstitial': ['interstitial'],
                   'bulk': ['bulk'], 'subsites': ['subsites'], 'equivalentsites': ['equivalencesites'],
                   'equivalencesites': ['equivalencesites']}

        def parse(arg, table):
            out = []
            for x in flatten(arg if isinstance(arg, list) else [arg]):
                if isinstance(x, (int, numpy.integer)) and not isinstance(x, bool):
                    out.append(int(x))
                elif x in table:
                    out.extend(table[x])
                else:
                    raise ValueError(str(x) + " must be a (list of) keyword(s) from " + str(list(table.keys()))
                                     + " or a (list of) integer(s) (i.e. site labels).")
            return out

        inclusions = parse(include, incl_kw)
        exclusions = parse(exclude, excl_kw)
        if self._n_labels is None:
            return numpy.array([])
        all_labels = list(range(self._n_labels))
        selected = []
        for inc in inclusions:
            if inc == 'default':
                exclusions.append('subsites')
                selected.extend(all_labels)
            elif inc == 'all':
                selected.extend(all_labels)
            elif inc == 'sites':
                exclusions.extend(['bulk', 'interstitial'])
                exclusions = [e for e in exclusions if e != 'subsites']
                selected.extend(all_labels)
            elif inc in ('equivalencesites', 'subsites'):
                exclusions = [e for e in exclusions if e != inc]
            else:
                if inc not in all_labels:
                    raise ValueError('site label %d does not exist' % inc)
                selected.append(inc)
        selected = set(selected)
        for exc in set(exclusions):
            if exc == 'interstitial':
                selected.discard(SITELABEL['interstitial'])
            elif exc == 'bulk' and 'bulk_site' in self.P:
                selected.discard(self.P['bulk_site'])
        return numpy.array(sorted(selected), dtype=int)

    def stats(self, data=None):
        """Statistics for the density, excluding bulk, interstitial and subsites
        (hop/sitemap.py:1301-1339): rho_cut, rho_cut_bulk, N_sites, N_equivalence_sites, N_subsites,
        site_volume_avg/std, site_occupancy_rho_avg/std.  `data` (a dict) receives the per-site
        arrays site_volume, site_occupancy_rho_avg, site_occupancy_rho_std."""
        if self.site_properties is None:
            raise AttributeError('Stats require site_properties annotation.')
        nodes = self.site_labels(exclude=['bulk', 'interstitial', 'subsites'])
        sp = self.site_properties.take(nodes)
        stats = {}
        if 'threshold' in self.P and 'bulk_threshold' in self.P:
            stats['rho_cut'], stats['rho_cut_bulk'] = self.P['threshold'], self.P['bulk_threshold']
        else:
            if 'threshold' in self.P:        # the reference sets rho_cut before it trips over the missing bulk site
                stats['rho_cut'] = self.P['threshold']
            warnings.warn("No bulk site defined", category=MissingDataWarning)
        # the key names are the contract (hop/sitemap.py:1318-1332); equivalence sites are not part of this package
        stats.update(N_sites=len(nodes), N_equivalence_sites=0, N_subsites=0,
                     site_volume_avg=numpy.average(sp.volume), site_volume_std=numpy.std(sp.volume),
                     site_occupancy_rho_avg=numpy.average(sp.occupancy_avg),
                     site_occupancy_rho_std=numpy.std(sp.occupancy_avg))
        if data is not None:
            try:
                data.update(site_volume=sp.volume, site_occupancy_rho_avg=sp.occupancy_avg,
                            site_occupancy_rho_std=sp.occupancy_std)
            except (TypeError, AttributeError):
                pass
        return stats

    def _draw_map_from_sites(self):
        """Repaint the map from the (possibly edited) site list (hop/sitemap.py:1574-1587)."""
        self._own = None
        self._require_device_sites()
        from . import kernels as K

        self._map16 = K.labels_to_map16(self._own)
        self._map_host = None

    def masked_density(self, density, site_labels):
        """Portion of `density` on the given sites; everything else zero."""
        m = numpy.isin(self.map, numpy.asarray(asiterable(site_labels)))
        return numpy.where(m, density, 0.0)

    def __repr__(self):
        features = [str(tuple(self.grid.shape)) + ' bins']
        try:
            features.append('default filename "' + str(self.filename()) + '"')
        except ValueError:
            pass
        if self._n_labels is not None:
            features.append(str(self._n_labels - 1) + ' sites')
        return '<hop_b200.sitemap.Density density with ' + ', '.join(features) + '>'

"""Runs the UNMODIFIED hop sources from /root/reference under Python 3 -- test infrastructure only.

hop 0.4.0-dev is Python 2.7 code on top of MDAnalysis 0.18 and cannot be imported here as it is
(SURVEY.md 8c).  This loader reads the reference's own files *where they lie* (nothing is copied into
the repository), translates them in memory and registers them as the package ``hop`` so that
``tests/golden/make_reference_golden.py`` can call the reference's real methods
(``DensityCollector.collect``, ``Density.map_sites``, ``site_insert_bulk``,
``HoppingTrajectory._coord2hop``, ``TransportNetwork._analyze_hops`` ...) and freeze their outputs.

The translation is mechanical and limited to what Python 3 / numpy 2 refuse or changed:

* syntax: ``except X, e`` -> ``except X as e``; ``print`` statements -> calls; ``<>`` -> ``!=``;
  ``raise E, msg`` -> ``raise E(msg)``; ``d.has_key(k)`` -> ``k in d``; ``it.next()`` -> ``next(it)``;
  ``from itertools import izip`` -> the lazy builtin ``zip``;
* Python 2 builtins, injected into each module's globals: ``zip``/``map``/``filter`` return lists,
  ``xrange``, ``unicode``, ``basestring``, ``long``, ``reduce``; ``cPickle`` is ``pickle``;
* classic division: in modules without ``from __future__ import division`` every ``a / b`` becomes
  ``_py2div(a, b)`` (floor division when both operands are integers, like Python 2) by an AST pass;
* numpy API drift (``NUMPY_PATCHES`` below, each an exact one-line replacement):
  ``histogramdd(normed=False)`` (keyword removed in numpy 1.24; False was the default) and indexing
  with a *list* of slices (an error since numpy 1.23; the tuple form is what it always meant);
  ``numpy.float_`` (alias of float64, removed in numpy 2.0).

MDAnalysis, gridData and Bio are replaced by small in-memory fakes (``install_fakes``): a Universe
over a float32 ``[F][N][3]`` array, the unit tables (MDAnalysis 0.18 values, the same constants as
``hop_b200/constants.py`` -- so unit conversion is *not* independently pinned, thresholds in the
golden cases are given in Angstrom^-3), a DCD ``Writer``/reader pair that keeps frames in memory.

One semantic caveat remains and is recorded in the golden files: ``_coord2hop``'s rounded top-edge
comparison (hop/trajectory.py:435-441) compares a float32 array with a float64 scalar, which numpy 1.x
evaluated in float32 and numpy 2 (NEP 50) evaluates in float64.  The golden cases keep every
coordinate away from the top edge, where both agree; the numpy-1.x rule is covered by the oracle's
own known-answer tests.
"""
from __future__ import annotations

import ast
import builtins
import functools
import os
import pickle
import re
import sys
import types

import numpy

REFERENCE_ROOT = os.environ.get("HOP_REFERENCE_ROOT", "/root/reference")
MODULES = ["exceptions", "constants", "utilities", "sitemap", "density", "trajectory", "graph"]

#: exact replacements for numpy API drift: (module, old, new)
NUMPY_PATCHES = [
    ("density", "range=self.arange, normed=False)", "range=self.arange)"),
    ("trajectory", "core = D*[slice(1,-1)]", "core = tuple(D*[slice(1,-1)])"),
    ("*", "numpy.float_", "numpy.float64"),          # alias removed in numpy 2.0
    ("*", "numpy.rank(", "numpy.ndim("),              # removed in numpy 1.18 (same function)
]


def available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "hop"))


# ---------------------------------------------------------------------------------------------
# source translation
# ---------------------------------------------------------------------------------------------

def _fix_print(line):
    m = re.match(r"^(\s*)((?:def \w+\(\): )?)print\b(.*)$", line)
    if not m or m.group(3).lstrip().startswith("("):
        return line
    indent, prefix, rest = m.groups()
    rest = rest.strip()
    if rest.startswith(">>"):
        target, _, rest = rest[2:].partition(",")
        return "%s%sprint(%s, file=%s)" % (indent, prefix, rest.strip(), target.strip())
    if rest.endswith(","):
        return "%s%sprint(%s end=' ')" % (indent, prefix, rest)
    return "%s%sprint(%s)" % (indent, prefix, rest)


def translate_source(name, src):
    for mod, old, new in NUMPY_PATCHES:
        if mod in (name, "*"):
            src = src.replace(old, new)
    src = re.sub(r"except\s+([\w\.]+)\s*,\s*(\w+)\s*:", r"except \1 as \2:", src)
    src = re.sub(r"except\s+\(([^)]*)\)\s*,\s*(\w+)\s*:", r"except (\1) as \2:", src)
    src = src.replace("<>", "!=")
    src = re.sub(r"raise\s+(\w+)\s*,\s*(.+)$", r"raise \1(\2)", src, flags=re.M)
    src = re.sub(r"([\w\.]+(?:\[[^\]]*\])?)\.has_key\(([^()]*)\)", r"(\2 in \1)", src)
    src = re.sub(r"\b([\w\.]+)\.next\(\)", r"next(\1)", src)
    src = re.sub(r"^from itertools import izip\s*$", "izip = __builtins_zip__", src, flags=re.M)
    src = src.replace(".iteritems()", ".items()").replace(".itervalues()", ".values()")
    src = src.replace(".iterkeys()", ".keys()")
    lines, out, i = src.split("\n"), [], 0
    while i < len(lines):
        line = lines[i]
        if re.match(r"^\s*(?:def \w+\(\): )?print\b\s*[^(\s]", line):
            # a print statement may continue over several lines (backslash or open brackets)
            def open_brackets(t):
                return sum(t.count(c) for c in "([{") - sum(t.count(c) for c in ")]}")
            n_joined = 0
            while (line.rstrip().endswith("\\") or open_brackets(line) > 0) and i + 1 < len(lines):
                i += 1
                n_joined += 1
                line = line.rstrip().rstrip("\\") + " " + lines[i].strip()
            out.append(_fix_print(line))
            out.extend([""] * n_joined)          # keep line numbers
        else:
            out.append(line)
        i += 1
    return "\n".join(out)


def _py2div(a, b):
    """Python 2 classic division: floor for integer operands (also numpy integers/arrays)."""
    def is_int(x):
        if isinstance(x, (bool, int, numpy.integer)):
            return True
        return isinstance(x, numpy.ndarray) and x.dtype.kind in "iub"
    if is_int(a) and is_int(b):
        return a // b
    return a / b


class _ClassicDivision(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()), args=[node.left, node.right],
                         keywords=[]), node)
        return node

    def visit_AugAssign(self, node):
        self.generic_visit(node)
        return node            # in-place `/=` on float arrays (density.py:136): same in 2 and 3


def compile_module(name, path):
    with open(path) as fh:
        src = translate_source(name, fh.read())
    tree = ast.parse(src, filename=path)
    future_division = any(isinstance(n, ast.ImportFrom) and n.module == "__future__"
                          and any(a.name == "division" for a in n.names) for n in tree.body)
    if not future_division:
        tree = ast.fix_missing_locations(_ClassicDivision().visit(tree))
    return compile(tree, path, "exec")


def _py2_globals():
    def lzip(*a):
        return list(builtins.zip(*a))

    def lmap(f, *a):
        if f is None:
            return lzip(*a) if len(a) > 1 else list(a[0])
        return list(builtins.map(f, *a))

    def lfilter(f, seq):
        return list(builtins.filter(f, seq))

    return dict(zip=lzip, map=lmap, filter=lfilter, xrange=range, unicode=str, basestring=str,
                long=int, reduce=functools.reduce, _py2div=_py2div, __builtins_zip__=builtins.zip)


# ---------------------------------------------------------------------------------------------
# fakes for MDAnalysis / gridData / Bio
# ---------------------------------------------------------------------------------------------

class FakeTimestep(object):
    """MDAnalysis.coordinates.base.Timestep: float32 _pos, _x/_y/_z views, 0-based frame."""

    def __init__(self, n_atoms):
        self.n_atoms = n_atoms
        self._pos = numpy.zeros((n_atoms, 3), dtype=numpy.float32)
        self.frame = 0
        self.dimensions = numpy.zeros(6, dtype=numpy.float32)

    _x = property(lambda self: self._pos[:, 0])
    _y = property(lambda self: self._pos[:, 1])
    _z = property(lambda self: self._pos[:, 2])
    positions = property(lambda self: self._pos)


class FakeReader(object):
    """In-memory trajectory reader over xyz float32 [F][N][3] (iteration, slicing, next, rewind)."""

    def __init__(self, xyz, dt=1.0, filename="mem.dcd"):
        self.xyz = numpy.ascontiguousarray(xyz, dtype=numpy.float32)
        self.n_frames, self.n_atoms = self.xyz.shape[:2]
        self.dt = dt
        self.totaltime = (self.n_frames - 1) * dt
        self.filename = filename
        self.ts = FakeTimestep(self.n_atoms)
        self._load(0)

    def _load(self, i):
        self.ts._pos[:] = self.xyz[i]
        self.ts.frame = i
        return self.ts

    def rewind(self):
        return self._load(0)

    def next(self):
        if self.ts.frame + 1 >= self.n_frames:
            raise StopIteration
        return self._load(self.ts.frame + 1)

    __next__ = next

    def __iter__(self):
        # Stays on the last frame when the iteration ends: `_analyze_hops` reads `self.traj.ts.frame`
        # after its loop as "the last frame" (hop/graph.py:259-264).  (Some MDAnalysis releases rewind
        # here; the drivers rewind explicitly wherever hop's callers start from frame 0.)
        for i in range(self.n_frames):
            yield self._load(i)

    def __getitem__(self, item):
        if isinstance(item, slice):
            def gen():
                for i in range(*item.indices(self.n_frames)):
                    yield self._load(i)
            return gen()
        return self._load(int(item))

    def __len__(self):
        return self.n_frames


class FakeAtomGroup(object):
    def __init__(self, universe, index):
        self.universe = universe
        self.index = numpy.asarray(index, dtype=numpy.int64)
        self.n_atoms = len(self.index)
        self.charges = numpy.zeros(self.n_atoms)
        self.masses = numpy.ones(self.n_atoms)
        self.types = numpy.array(["O"] * self.n_atoms)

    atoms = property(lambda self: self)
    positions = property(lambda self: self.universe.trajectory.ts._pos[self.index])

    def __len__(self):
        return self.n_atoms

    def __iter__(self):
        solute = set(self.universe.selections.get("protein", ()))
        for i in self.index:
            i = int(i)
            yield types.SimpleNamespace(
                index=i, segid="SYSTEM", resid=i + 1, resname="ALA" if i in solute else "TIP3",
                name="CA" if i in solute else "OH2", type="CT1" if i in solute else "OT",
                charge=0.0, mass=12.011 if i in solute else 15.9994)


class FakeUniverse(object):
    """Universe over an in-memory trajectory; `selections` maps selection strings to index arrays."""
    _files = {}          # "file name" -> FakeReader, filled by FakeWriter

    def __init__(self, *args, **kwargs):
        if args and isinstance(args[0], numpy.ndarray):
            self.trajectory = FakeReader(args[0], dt=kwargs.get("dt", 1.0))
            self.filename = "mem.psf"
        else:                                    # Universe(hoppsf, hopdcd) of hop/trajectory.py:201
            self.filename, dcd = args[0], args[1]
            self.trajectory = self._files[dcd]
            self.trajectory.rewind()
        n = self.trajectory.n_atoms
        self.selections = dict(kwargs.get("selections", {}))
        self.selections.setdefault("all", numpy.arange(n))
        self.atoms = FakeAtomGroup(self, numpy.arange(n))

    def select_atoms(self, sel):
        return FakeAtomGroup(self, self.selections[sel])


class FakeWriter(object):
    """MDAnalysis.Writer(dcdname, n_atoms=, dt=, remarks=): collects frames, registers a reader."""

    def __init__(self, filename, n_atoms=None, dt=1.0, remarks="", **kwargs):
        self.filename, self.dt, self.frames = filename, dt, []

    def write_next_timestep(self, ts):
        self.frames.append(ts._pos.copy())

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        FakeUniverse._files[self.filename] = FakeReader(numpy.array(self.frames), dt=self.dt,
                                                        filename=self.filename)
        return False


class FakeProgressMeter(object):
    def __init__(self, *args, **kwargs):
        pass

    def echo(self, *args, **kwargs):
        pass


def _notwithin_coordinates_factory(universe, sel1, sel2, cutoff, not_within=True, use_kdtree=True):
    """MDAnalysis.analysis.density.notwithin_coordinates_factory: coordinates of sel1 atoms that
    have no sel2 atom within cutoff (brute-force distances; the selection itself is MDAnalysis code,
    not hop code)."""
    solvent, solute = universe.select_atoms(sel1), universe.select_atoms(sel2)

    def notwithin_coordinates():
        a = solvent.positions.astype(numpy.float64)
        b = solute.positions.astype(numpy.float64)
        d2 = ((a[:, None, :] - b[None, :, :]) ** 2).sum(-1)
        near = (d2 <= float(cutoff) ** 2).any(axis=1)
        return solvent.positions[near != not_within]
    return notwithin_coordinates


def install_fakes():
    def module(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    N_A = 6.02214179e+23
    water = {"exp": 0.997, "SPC": 0.985, "TIP3P": 1.002, "TIP4P": 1.001, "MolarMass": 18.016}
    length = {"Angstrom": 1.0, "A": 1.0, "angstrom": 1.0, "nm": 1.0 / 10, "nanometer": 1.0 / 10}
    dens = {"Angstrom^{-3}": 1 / 1.0, "A^{-3}": 1 / 1.0, "nm^{-3}": 1 / 1e-3,
            "Molar": 1 / (1e-27 * N_A),
            "SPC": 1 / (1e-24 * N_A * water["SPC"] / water["MolarMass"]),
            "TIP3P": 1 / (1e-24 * N_A * water["TIP3P"] / water["MolarMass"]),
            "TIP4P": 1 / (1e-24 * N_A * water["TIP4P"] / water["MolarMass"]),
            "water": 1 / (1e-24 * N_A * water["exp"] / water["MolarMass"])}
    time = {"ps": 1.0, "ns": 1e-3}
    conversion_factor = {"length": length, "density": dens, "time": time}

    def get_conversion_factor(unit_type, u1, u2):
        return conversion_factor[unit_type][u2] / conversion_factor[unit_type][u1]

    def convert(x, u1, u2):
        for t, table in conversion_factor.items():
            if u1 in table:
                return x * get_conversion_factor(t, u1, u2)
        raise ValueError(u1)

    units = module("MDAnalysis.units", constants={"N_Avogadro": N_A}, water=water,
                   conversion_factor=conversion_factor, get_conversion_factor=get_conversion_factor,
                   convert=convert, densityUnit_factor=dens, lengthUnit_factor=length)
    log = module("MDAnalysis.lib.log", ProgressMeter=FakeProgressMeter)
    lib = module("MDAnalysis.lib", log=log)
    base = module("MDAnalysis.coordinates.base", Timestep=FakeTimestep)
    dcd = module("MDAnalysis.coordinates.DCD", DCDReader=FakeReader)
    coordinates = module("MDAnalysis.coordinates", base=base, DCD=dcd)
    topologyattrs = module("MDAnalysis.core.topologyattrs")
    core = module("MDAnalysis.core", flags={"length_unit": "Angstrom"}, topologyattrs=topologyattrs)
    adens = module("MDAnalysis.analysis.density",
                   notwithin_coordinates_factory=_notwithin_coordinates_factory,
                   BfactorDensityCreator=object)
    analysis = module("MDAnalysis.analysis", density=adens)
    module("MDAnalysis", units=units, lib=lib, coordinates=coordinates, core=core, analysis=analysis,
           Universe=FakeUniverse, Writer=FakeWriter, as_Universe=lambda u, *a, **k: u,
           __version__="0.18-fake")
    module("gridData", OpenDX=types.SimpleNamespace())
    module("Bio.PDB")
    module("Bio", PDB=sys.modules["Bio.PDB"])
    # cPickle of Python 2: highest protocol 2
    module("cPickle", dump=pickle.dump, dumps=pickle.dumps, load=pickle.load, loads=pickle.loads,
           HIGHEST_PROTOCOL=2, PickleError=pickle.PickleError)


# ---------------------------------------------------------------------------------------------
# loading
# ---------------------------------------------------------------------------------------------

def unload_reference():
    """Removes the translated reference and the fakes from sys.modules again."""
    ours = getattr(sys.modules.get("hop"), "__reference_py3__", False)
    for m in list(sys.modules):
        root = m.split(".")[0]
        if root in ("MDAnalysis", "gridData", "Bio") or m == "cPickle" or (root == "hop" and ours):
            del sys.modules[m]


def load_reference():
    """Returns the package object `hop` built from the translated reference sources."""
    if "hop" in sys.modules and getattr(sys.modules["hop"], "__reference_py3__", False):
        return sys.modules["hop"]
    if not available():
        raise RuntimeError("reference sources not found under %s" % REFERENCE_ROOT)
    install_fakes()
    pkg_dir = os.path.join(REFERENCE_ROOT, "hop")
    pkg = types.ModuleType("hop")
    pkg.__path__ = [pkg_dir]
    pkg.__package__ = "hop"
    pkg.__reference_py3__ = True
    sys.modules["hop"] = pkg
    for name in MODULES:
        mod = types.ModuleType("hop." + name)
        mod.__file__ = os.path.join(pkg_dir, name + ".py")
        mod.__package__ = "hop"
        mod.__dict__.update(_py2_globals())
        sys.modules["hop." + name] = mod
        setattr(pkg, name, mod)
        exec(compile_module(name, mod.__file__), mod.__dict__)
    return pkg


def load_siteanalysis():
    """hop.siteanalysis of the reference, compiled like the other modules.  Two things it expects from its
    surroundings are supplied here: `hop.MissingDataError` (the package __init__, which is not executed, exports
    it) and `numpy.NaN` (removed in numpy 2; an alias of numpy.nan)."""
    hop = load_reference()
    if hasattr(hop, "siteanalysis"):
        return hop.siteanalysis
    hop.MissingDataError = hop.exceptions.MissingDataError
    if not hasattr(numpy, "NaN"):
        numpy.NaN = numpy.nan
    name = "siteanalysis"
    mod = types.ModuleType("hop." + name)
    mod.__file__ = os.path.join(REFERENCE_ROOT, "hop", name + ".py")
    mod.__package__ = "hop"
    mod.__dict__.update(_py2_globals())
    sys.modules["hop." + name] = mod
    exec(compile_module(name, mod.__file__), mod.__dict__)
    hop.siteanalysis = mod
    return mod


class LegacySelection(object):
    """An AtomGroup as hop.siteanalysis.Collection.__init__ uses it (hop/siteanalysis.py:943-989): the MDAnalysis
    API from before release 0.11 -- `selection.indices()` is a method, atoms carry `.segment.name` and
    `.residue.id`, `selection.atoms[0].universe` is the Universe."""

    def __init__(self, group):
        self._group = group
        self.universe = group.universe

    atoms = property(lambda self: self)

    def indices(self):
        return numpy.asarray(self._group.index)

    def __len__(self):
        return len(self._group)

    def __iter__(self):
        for a in self._group:
            yield types.SimpleNamespace(universe=self._group.universe, segment=types.SimpleNamespace(name=a.segid),
                                        residue=types.SimpleNamespace(id=a.resid))

    def __getitem__(self, i):
        return list(self)[i]

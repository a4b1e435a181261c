"""Support code on the hot path: grid sizing, the pickle-based Saveable, small helpers.

Mirrors the parts of hop/utilities.py that the four pipeline stages use
(fixedwidth_bins 373-389, Saveable 119-261, filename_function 69-106, DefaultDict 754-760).
"""
from __future__ import annotations

import os
import pickle
import warnings

import numpy

from .exceptions import MissingDataWarning


def fixedwidth_bins(delta, xmin, xmax):
    """Return bins of width delta that cover xmin,xmax (or a larger range).

    dict = fixedwidth_bins(delta,xmin,xmax); keys 'Nbins', 'delta', 'min', 'max'
    (hop/utilities.py:373-389).
    """
    if not numpy.all(xmin < xmax):
        raise ValueError('Boundaries are not sane: should be xmin < xmax.')
    _delta = numpy.asarray(delta, dtype=numpy.float64)
    _xmin = numpy.asarray(xmin, dtype=numpy.float64)
    _xmax = numpy.asarray(xmax, dtype=numpy.float64)
    _length = _xmax - _xmin
    N = numpy.ceil(_length / _delta).astype(numpy.int_)
    dx = 0.5 * (N * _delta - _length)
    return {'Nbins': N, 'delta': _delta, 'min': _xmin - dx, 'max': _xmax + dx}


def filename_function(self, filename=None, ext=None, set_default=False, use_my_ext=False):
    """Supply a file name for the object (hop/utilities.py:69-106)."""
    if filename is None:
        if not hasattr(self, '_filename'):
            self._filename = None
        if self._filename:
            filename = self._filename
        else:
            raise ValueError("A file name is required because no default file name was defined.")
        my_ext = None
    else:
        filename, my_ext = os.path.splitext(filename)
        if set_default:
            self._filename = filename
    if my_ext and use_my_ext:
        ext = my_ext
    if ext is not None:
        if ext.startswith('.'):
            ext = ext[1:]
        filename = filename + '.' + ext
    return filename


class _HopUnpickler(pickle.Unpickler):
    """Reads pickles written by hop itself: classes of the package ``hop`` (hop.sitemap.Density,
    hop.utilities.DefaultDict, hop.graph.HoppingGraph, hop.graph.fitExp2, ...) are resolved to their
    counterparts here, so that the files load where hop is not installed."""

    def find_class(self, module, name):
        if module == 'hop' or module.startswith('hop.'):
            module = 'hop_b200' + module[3:]
        return super(_HopUnpickler, self).find_class(module, name)


def load_pickle(fh):
    """pickle.load that also accepts files saved by hop (Python-2 string encoding included)."""
    return _HopUnpickler(fh, encoding='latin1').load()


class DefaultDict(dict):
    """Dictionary based on defaults and updated with keys/values from user."""

    def __init__(self, defaultdict, userdict=None, **kwargs):
        super(DefaultDict, self).__init__(**defaultdict)
        if userdict is not None:
            self.update(userdict)
        self.update(kwargs)


class Saveable(object):
    """Baseclass that supports save()ing and load()ing of the attributes named in
    ``_saved_attributes`` as one pickle (hop/utilities.py:119-261)."""

    _saved_attributes = []
    _merge_attributes = []
    _excluded_attributes = []

    def __init__(self, *args, **kwargs):
        kwargs.setdefault('filename', None)
        super(Saveable, self).__init__()
        if kwargs['filename'] is not None:
            self.load(kwargs['filename'], 'pickle')

    # The pickle SCHEMA is the contract with hop (a dict of the attributes named in _saved_attributes, or of
    # everything outside _excluded_attributes for 'all'; hop/utilities.py:154-261); files written by either
    # package load in the other.

    def _attribute_names(self, available):
        """The attribute names that take part in save / load, given the names that exist."""
        if self._saved_attributes == 'all':
            return [name for name in available if name not in self._excluded_attributes]
        return list(self._saved_attributes)

    def __getstate__(self):
        if not self._saved_attributes:
            warnings.warn("No data saved: the class declared empty '_saved_attributes'.")
            return False
        state = {}
        for name in self._attribute_names(self.__dict__.keys()):
            try:
                state[name] = self._get_saved_attribute(name)
            except (KeyError, AttributeError):
                warnings.warn("Attribute '%s' has not been computed and will not be saved." % name,
                              category=MissingDataWarning)
        return state

    def _get_saved_attribute(self, attr):
        return self.__dict__[attr]

    def __setstate__(self, data):
        self.__dict__.update(data)

    filename = filename_function

    def save(self, filename=None):
        """Save class to a pickled file."""
        target = self.filename(filename, 'pickle', set_default=True)
        with open(target, 'wb') as fh:
            pickle.dump(self, fh, pickle.HIGHEST_PROTOCOL)

    def load(self, filename=None, merge=False):
        """Reinstantiate class from a pickled file (produced with save())."""
        source = self.filename(filename, 'pickle', set_default=True)
        with open(source, 'rb') as fh:
            stored = load_pickle(fh)
        if isinstance(stored, Saveable):
            # attributes may live behind properties (device mirrors): ask the object, not its __dict__
            data = stored.__getstate__()
        elif hasattr(stored, '__dict__'):
            data = stored.__dict__
        else:
            # a bare dict from the time before hop pickled instances; 'parameters' was renamed to 'P'
            warnings.warn("Loading an old-style save file.", category=DeprecationWarning)
            data = stored
            if 'parameters' in data:
                data['P'] = data['parameters']
        if not self._saved_attributes:
            warnings.warn("No data loaded: the object declared empty '_saved_attributes'.",
                          category=MissingDataWarning)
            return False
        for name in self._attribute_names(data.keys()):
            if name not in data:
                warnings.warn("Expected attribute '%s' was not found in saved file '%s'." % (name, filename),
                              category=MissingDataWarning)
            elif merge and name in self._merge_attributes:
                self.__dict__[name].update(data[name])       # dict-valued attributes only
            else:
                self._set_loaded_attribute(name, data[name])
        return True

    def _set_loaded_attribute(self, attr, value):
        self.__dict__[attr] = value


def iterable(obj):
    """Returns True if obj can be iterated over and is NOT a string."""
    if isinstance(obj, str):
        return False
    if hasattr(obj, 'next') or hasattr(obj, '__next__'):
        return True
    try:
        len(obj)
    except TypeError:
        return False
    return True


def asiterable(obj):
    """Return an object that is an iterable: object itself or wrapped in a list."""
    if not iterable(obj):
        obj = [obj]
    return obj


def flatten(l):
    """Flatten nested lists/tuples into a flat list."""
    out = []
    for x in l:
        if isinstance(x, (list, tuple)):
            out.extend(flatten(x))
        else:
            out.append(x)
    return out


def easy_load(names, baseclass, keymethod):
    """Instantiate a class either from an existing instance or a pickled file
    (hop/utilities.py:263-296): objects that implement `keymethod` are taken as they are, anything
    else is treated as a file name for ``baseclass(filename=name)``.  A single name gives a single
    instance, a list gives a list."""
    def load(name):
        return name if hasattr(name, keymethod) else baseclass(filename=name)

    if iterable(names) and not hasattr(names, keymethod):
        return [load(x) for x in names]
    return load(names)

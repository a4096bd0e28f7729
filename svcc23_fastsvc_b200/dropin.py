"""Merging the `harana` drop-in namespace of this repo with the reference's own `harana` package.

`harana` is a namespace package in the reference (no ``harana/__init__.py``), but ``harana.models``,
``harana.utils`` and ``harana.layers`` are regular packages, so when this repo precedes the reference on
``sys.path`` our shim packages shadow the reference's completely.  The scripts the drop-in has to satisfy
need names that live only in the reference (``getattr(harana.models, "MelGANMultiScaleDiscriminator")`` at
train_fastsvc.py:705, ``from harana.utils import read_hdf5`` at :38, ``harana.losses`` importing
``make_non_pad_mask`` from ``harana.utils``), so each shim package

  1. extends its ``__path__`` over the later ``sys.path`` entries (``pkgutil.extend_path``), which makes the
     reference's sibling modules (``harana.models.tacotron2``, ``harana.utils.utils`` ...) importable;
  2. loads the reference module its own same-named file shadows (``fastsvc.py``, ``features.py`` ...) under
     ``<package>._reference_<name>`` and re-exports its public names, then overrides the hot-path classes
     with the ones backed by libfsvc.so.

When the reference is not importable (the GPU box), the shims expose just the hot-path classes.
"""

import importlib
import importlib.util
import os
import sys


def _reference_dirs(pkg_path, own_file):
    own = os.path.dirname(os.path.abspath(own_file))
    return [p for p in pkg_path if os.path.abspath(p) != own]


def public_names(mod):
    names = getattr(mod, "__all__", None)
    if names is None:
        names = [n for n in vars(mod) if not n.startswith("_")]
    return list(names)


def load_shadowed(pkg_name, own_file, basename):
    """The reference's ``<pkg_name>.<basename>`` that our same-named file shadows, or None."""
    pkg = sys.modules[pkg_name]
    for d in _reference_dirs(pkg.__path__, own_file):
        path = os.path.join(d, basename + ".py")
        if os.path.isfile(path):
            name = f"{pkg_name}._reference_{basename}"
            if name in sys.modules:
                return sys.modules[name]
            spec = importlib.util.spec_from_file_location(name, path)
            mod = importlib.util.module_from_spec(spec)
            sys.modules[name] = mod
            try:
                spec.loader.exec_module(mod)
            except BaseException:
                del sys.modules[name]
                raise
            return mod
    return None


def adopt_shadowed(module_globals, pkg_name, own_file, basename, ours):
    """Re-export the shadowed reference module's public names into ``module_globals`` (a shim module),
    keeping ``ours`` (names already defined there); returns the resulting ``__all__``."""
    names = list(ours)
    ref = load_shadowed(pkg_name, own_file, basename)
    if ref is not None:
        for n in public_names(ref):
            if n not in ours:
                module_globals[n] = getattr(ref, n)
                names.append(n)
    return names


def import_siblings(module_globals, pkg_name, own_file):
    """Import every reference module of the package that this repo does not shadow and star-export it
    (what the reference's ``__init__`` does with ``from .x import *``)."""
    own = os.path.dirname(os.path.abspath(own_file))
    mine = {os.path.splitext(f)[0] for f in os.listdir(own)}
    pkg = sys.modules[pkg_name]
    for d in _reference_dirs(pkg.__path__, own_file):
        if not os.path.isdir(d):
            continue
        for f in sorted(os.listdir(d)):
            base, ext = os.path.splitext(f)
            if ext != ".py" or base.startswith("_") or base in mine:
                continue
            mod = importlib.import_module(f"{pkg_name}.{base}")
            for n in public_names(mod):
                module_globals.setdefault(n, getattr(mod, n))

"""TEST INFRASTRUCTURE ONLY -- runtime compatibility shim for importing the *unmodified* reference.

The reference (mars-project/mars) pins ``pandas<2`` / ``numpy<2``; this image has pandas 3 / numpy 2 and
no ``tornado``.  Importing this module *before* ``mars`` installs the aliases the reference touches on
the groupby-aggregation path so that its own ``execute(ctx, op)`` code can be run here (build container
only -- ``/root/reference`` does not exist on the GPU box) to

* validate ``oracle/mars_groupby_oracle.py`` (the CPU restatement), and
* generate the golden fixtures under ``tests/golden/`` (``oracle/make_golden.py``).

Nothing in ``mars_b200/`` imports this file.  No reference source is copied: the reference is built from
a scratch copy by ``oracle/build_ref.sh`` and imported from there.
"""
from __future__ import annotations

import sys
import types
import warnings

import numpy as np
import pandas  # noqa: F401  (must be imported before the tornado stub is installed)
import pandas as pd

try:  # scipy / sklearn are optional for the path but must precede the stubs if present
    import scipy  # noqa: F401
    import sklearn  # noqa: F401
except Exception:  # pragma: no cover
    pass

warnings.filterwarnings("ignore")


def _alias(mod, name, value):
    if not hasattr(mod, name):
        try:
            setattr(mod, name, value)
        except Exception:  # pragma: no cover
            pass


def _install_numpy():
    for name, value in dict(
        unicode_=np.str_,
        float_=np.float64,
        complex_=np.complex128,
        cfloat=np.complex128,
        bool8=np.bool_,
        string_=np.bytes_,
        int0=np.intp,
        uint0=np.uintp,
        object0=np.object_,
        str0=np.str_,
        bytes0=np.bytes_,
        void0=np.void,
        longfloat=np.longdouble,
        clongfloat=np.clongdouble,
        longcomplex=np.clongdouble,
        singlecomplex=np.complex64,
        Inf=np.inf,
        NaN=np.nan,
        NAN=np.nan,
        PINF=np.inf,
        NINF=-np.inf,
        PZERO=0.0,
        NZERO=-0.0,
        Infinity=np.inf,
        infty=np.inf,
    ).items():
        _alias(np, name, value)
    try:
        from numpy.exceptions import AxisError, ComplexWarning
    except Exception:  # pragma: no cover
        AxisError = ComplexWarning = Exception
    _alias(np, "AxisError", AxisError)

    def find_common_type(array_types, scalar_types):
        return np.result_type(*array_types, *scalar_types)

    _alias(np, "find_common_type", find_common_type)
    try:
        import numpy.core.numerictypes as nt

        _alias(nt, "find_common_type", find_common_type)
        import numpy.core.numeric as nn

        _alias(nn, "ComplexWarning", ComplexWarning)
    except Exception:  # pragma: no cover
        pass
    if "numpy.compat" not in sys.modules:
        m = types.ModuleType("numpy.compat")
        m.basestring = str
        sys.modules["numpy.compat"] = m
        np.compat = m
    if "numpy.lib.index_tricks" not in sys.modules:
        m = types.ModuleType("numpy.lib.index_tricks")
        m.ndindex = np.ndindex
        sys.modules["numpy.lib.index_tricks"] = m


def _install_pandas():
    for cls in (pd.DataFrame, pd.Series):
        if not hasattr(cls, "_data"):
            cls._data = property(lambda self: self._mgr)
    try:
        import pandas._config.config as cf

        try:
            cf.get_option("mode.use_inf_as_na")
        except Exception:
            with cf.config_prefix("mode"):
                cf.register_option("use_inf_as_na", False, "compat shim", validator=cf.is_bool)
    except Exception:  # pragma: no cover
        pass
    from pandas.core.groupby import GroupBy

    if not hasattr(GroupBy, "grouper"):
        GroupBy.grouper = property(lambda self: self._grouper)


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return _Dummy()


class _StubModule(types.ModuleType):
    def __getattr__(self, item):
        if item.startswith("__") and item.endswith("__"):
            raise AttributeError(item)
        return type(item, (_Dummy,), {})


def _install_tornado():
    try:
        import tornado  # noqa: F401

        return
    except ImportError:
        pass
    names = [
        "tornado",
        "tornado.web",
        "tornado.ioloop",
        "tornado.httpclient",
        "tornado.httpserver",
        "tornado.netutil",
        "tornado.gen",
        "tornado.escape",
        "tornado.websocket",
        "tornado.simple_httpclient",
        "tornado.log",
        "tornado.platform",
        "tornado.platform.asyncio",
        "tornado.concurrent",
    ]
    for n in names:
        m = _StubModule(n)
        m.__path__ = []
        sys.modules[n] = m
    for n in names[1:]:
        parent, _, child = n.rpartition(".")
        setattr(sys.modules[parent], child, sys.modules[n])


def install():
    _install_numpy()
    _install_pandas()
    _install_tornado()


install()

"""Labelled arrays for simulation results.

The reference returns ``xarray`` objects from ``run_simulation``
(/root/reference/src/openlifu/sim/kwave_if.py:131-146) and post-processes them in
``plan/solution.py`` / ``plan/solution_analysis.py``.  When ``xarray`` is importable this
module re-exports it unchanged, so results ARE xarray objects.  When it is not (the build and
GPU images of this repo have no xarray), the classes below provide the part of the surface
that code path touches: ``DataArray``, ``Dataset``, ``Coordinates``,
``concat``, with name-based broadcasting, ``isel/sel/where/max/mean/sum/interp`` and writable
``.data`` views.
"""
from __future__ import annotations

import copy as _copy
import operator
from collections import OrderedDict
from typing import Any, Iterable, Mapping

import numpy as np

try:  # pragma: no cover - exercised only where xarray exists
    import xarray as _xr

    DataArray = _xr.DataArray
    Dataset = _xr.Dataset
    Coordinates = _xr.Coordinates
    concat = _xr.concat
    open_dataset = _xr.open_dataset
    HAVE_XARRAY = True
except Exception:  # noqa: BLE001
    HAVE_XARRAY = False


if not HAVE_XARRAY:

    def _as_index_array(v):
        if isinstance(v, DataArray):
            return v
        return np.asarray(v)

    class _Indexes:
        """Tiny stand-in for the mapping returned by ``.coords``."""

    class Coordinates(Mapping):
        """Ordered mapping ``dim name -> 1-D DataArray`` (index coordinates only, plus scalars)."""

        def __init__(self, coords: Mapping | None = None):
            self._vars: "OrderedDict[str, DataArray]" = OrderedDict()
            if coords is not None:
                for name, val in coords.items():
                    self[name] = val

        # -- mapping protocol
        def __getitem__(self, key):
            return self._vars[key]

        def __setitem__(self, key, val):
            self._vars[key] = _coord_var(key, val)

        def __iter__(self):
            return iter(self._vars)

        def __len__(self):
            return len(self._vars)

        def __contains__(self, key):
            return key in self._vars

        def __repr__(self):
            rows = [f"  * {k} ({','.join(v.dims)}) {v.data.dtype} size {v.size}" for k, v in self._vars.items()]
            return "Coordinates:\n" + "\n".join(rows)

        @property
        def dims(self):
            return tuple(k for k, v in self._vars.items() if v.dims == (k,))

        @property
        def sizes(self):
            return {k: v.size for k, v in self._vars.items() if v.dims == (k,)}

        def copy(self, deep=True):
            out = Coordinates()
            for k, v in self._vars.items():
                out._vars[k] = v.copy(deep=deep)
            return out

        def to_dict(self):
            return dict(self._vars)

    def _coord_var(name, val) -> "DataArray":
        if isinstance(val, DataArray):
            if val.ndim == 0:
                return DataArray(val.data, dims=(), attrs=dict(val.attrs), name=name, _plain=True)
            return DataArray(val.data, dims=val.dims, attrs=dict(val.attrs), name=name, _plain=True)
        if isinstance(val, tuple) and len(val) in (2, 3) and isinstance(val[0], (str, tuple, list)):
            dims = (val[0],) if isinstance(val[0], str) else tuple(val[0])
            attrs = dict(val[2]) if len(val) == 3 else {}
            return DataArray(np.asarray(val[1]), dims=dims, attrs=attrs, name=name, _plain=True)
        arr = np.asarray(val)
        dims = (name,) if arr.ndim == 1 else ()
        return DataArray(arr, dims=dims, name=name, _plain=True)

    def _binary(op, reflexive=False):
        def f(self, other):
            return self._binary_op(other, op, reflexive)
        return f

    class DataArray:
        __array_priority__ = 60

        def __init__(self, data=np.nan, coords=None, dims=None, name=None, attrs=None, _plain=False):
            data = np.asarray(data) if not isinstance(data, np.ndarray) else data
            if dims is None:
                if coords is not None and not _plain:
                    cd = list(coords.dims) if isinstance(coords, Coordinates) else [
                        k for k, v in coords.items() if np.ndim(v if not isinstance(v, tuple) else v[1]) == 1]
                    dims = tuple(cd[: data.ndim]) if len(cd) >= data.ndim else tuple(f"dim_{i}" for i in range(data.ndim))
                else:
                    dims = tuple(f"dim_{i}" for i in range(data.ndim))
            if isinstance(dims, str):
                dims = (dims,)
            dims = tuple(dims)
            if len(dims) != data.ndim:
                raise ValueError(f"different number of dimensions on data ({data.ndim}) and dims ({dims})")
            self._data = data
            self.dims = dims
            self.name = name
            self.attrs = dict(attrs) if attrs else {}
            self._coords = Coordinates()
            if coords is not None and not _plain:
                items = coords.items()
                for k, v in items:
                    cv = _coord_var(k, v)
                    if all(d in dims for d in cv.dims):
                        for d, n in zip(cv.dims, cv.shape):
                            if n != data.shape[dims.index(d)]:
                                raise ValueError(f"conflicting sizes for dimension {d!r}")
                        self._coords._vars[k] = cv

        # -- basic properties
        @property
        def data(self):
            return self._data

        @data.setter
        def data(self, value):
            value = np.asarray(value)
            if value.shape != self._data.shape:
                raise ValueError("replacement data must match the shape")
            self._data = value

        values = data

        @property
        def coords(self):
            return self._coords

        @property
        def shape(self):
            return self._data.shape

        @property
        def ndim(self):
            return self._data.ndim

        @property
        def size(self):
            return int(self._data.size)

        @property
        def dtype(self):
            return self._data.dtype

        @property
        def sizes(self):
            return dict(zip(self.dims, self._data.shape))

        def __len__(self):
            return self._data.shape[0]

        def __array__(self, dtype=None, copy=None):
            return np.asarray(self._data, dtype=dtype)

        def to_numpy(self):
            return np.asarray(self._data)

        def item(self):
            return self._data.item()

        def __float__(self):
            return float(self._data)

        def __int__(self):
            return int(self._data)

        def __bool__(self):
            return bool(self._data)

        def __repr__(self):
            return f"<DataArray {self.name!r} {self.sizes} {self._data.dtype}>\n{self._data!r}"

        def __iter__(self):
            for i in range(len(self)):
                yield self[i]

        # -- construction helpers
        def _replace(self, data, dims=None, coords=None, keep_attrs=True, name="__same__"):
            out = DataArray.__new__(DataArray)
            out._data = data
            out.dims = self.dims if dims is None else tuple(dims)
            out.name = self.name if name == "__same__" else name
            out.attrs = dict(self.attrs) if keep_attrs else {}
            out._coords = Coordinates()
            src = self._coords if coords is None else coords
            for k, v in src._vars.items():
                if all(d in out.dims for d in v.dims):
                    out._coords._vars[k] = v
            return out

        def copy(self, deep=True, data=None):
            d = self._data if data is not None else (self._data.copy() if deep else self._data)
            out = self._replace(np.asarray(d), coords=self._coords.copy(deep=deep))
            out.attrs = _copy.deepcopy(self.attrs) if deep else dict(self.attrs)
            return out

        def __copy__(self):
            return self.copy(deep=False)

        def __deepcopy__(self, memo):
            return self.copy(deep=True)

        def astype(self, dtype):
            return self._replace(self._data.astype(dtype))

        def rename(self, name):
            out = self.copy(deep=False)
            out.name = name
            return out

        # -- indexing
        def __getitem__(self, key):
            if isinstance(key, str):
                return self._coords[key]
            if isinstance(key, DataArray):
                key = key.data
            if not isinstance(key, tuple):
                key = (key,)
            if any(k is Ellipsis for k in key):
                i = [k is Ellipsis for k in key].index(True)
                fill = (slice(None),) * (self.ndim - len(key) + 1)
                key = key[:i] + fill + key[i + 1:]
            key = key + (slice(None),) * (self.ndim - len(key))
            return self.isel({d: k for d, k in zip(self.dims, key)})

        def __setitem__(self, key, value):
            if isinstance(key, DataArray):
                key = key.data
            self._data[key] = value.data if isinstance(value, DataArray) else value

        def isel(self, indexers=None, **kw):
            idx = dict(indexers or {}, **kw)
            key = []
            new_dims = []
            for d in self.dims:
                k = idx.get(d, slice(None))
                if isinstance(k, DataArray):
                    k = k.data
                if isinstance(k, (list, np.ndarray)):
                    k = np.asarray(k)
                key.append(k)
                if not isinstance(k, (int, np.integer)):
                    new_dims.append(d)
            # apply one axis at a time so that array indexers stay orthogonal and ints/slices give views
            data = self._data
            ax = 0
            for k in key:
                sl = (slice(None),) * ax + (k,)
                data = data[sl]
                if not isinstance(k, (int, np.integer)):
                    ax += 1
            coords = Coordinates()
            for name, cv in self._coords._vars.items():
                if cv.dims == ():
                    coords._vars[name] = cv
                    continue
                d = cv.dims[0]
                k = key[self.dims.index(d)]
                sub = cv._data[k]
                if np.ndim(sub) == 0:
                    coords._vars[name] = DataArray(np.asarray(sub), dims=(), attrs=cv.attrs, name=name, _plain=True)
                else:
                    coords._vars[name] = DataArray(sub, dims=(d,), attrs=cv.attrs, name=name, _plain=True)
            out = DataArray.__new__(DataArray)
            out._data = data
            out.dims = tuple(new_dims)
            out.name = self.name
            out.attrs = dict(self.attrs)
            out._coords = Coordinates()
            for k2, v in coords._vars.items():
                if all(dd in out.dims for dd in v.dims):
                    out._coords._vars[k2] = v
            return out

        def sel(self, indexers=None, method=None, **kw):
            idx = dict(indexers or {}, **kw)
            pos = {}
            for d, lab in idx.items():
                c = self._coords[d].data
                if isinstance(lab, slice):
                    lo = -np.inf if lab.start is None else lab.start
                    hi = np.inf if lab.stop is None else lab.stop
                    pos[d] = np.flatnonzero((c >= lo) & (c <= hi))
                elif np.ndim(lab) == 0:
                    if method == "nearest":
                        pos[d] = int(np.argmin(np.abs(c - lab)))
                    else:
                        hit = np.flatnonzero(c == lab)
                        if hit.size == 0:
                            raise KeyError(f"{lab!r} not found in coordinate {d!r}")
                        pos[d] = int(hit[0])
                else:
                    lab = np.asarray(lab)
                    if method == "nearest":
                        pos[d] = np.array([int(np.argmin(np.abs(c - v))) for v in lab])
                    else:
                        pos[d] = np.array([int(np.flatnonzero(c == v)[0]) for v in lab])
            return self.isel(pos)

        # -- coordinates
        def assign_coords(self, coords=None, **kw):
            new = dict(coords or {}, **kw)
            out = self.copy(deep=False)
            out._coords = self._coords.copy(deep=False)
            for k, v in new.items():
                cv = _coord_var(k, v)
                if not all(d in out.dims for d in cv.dims):
                    raise ValueError(f"coordinate {k!r} has dims {cv.dims} not on the array {out.dims}")
                out._coords._vars[k] = cv
            return out

        def expand_dims(self, dim, axis=0):
            data = np.expand_dims(self._data, axis)
            dims = list(self.dims)
            dims.insert(axis, dim)
            out = self._replace(data, dims=dims)
            if dim in self._coords and self._coords[dim].ndim == 0:
                c = self._coords[dim]
                out._coords._vars[dim] = DataArray(np.asarray([c.data.item()]), dims=(dim,), attrs=c.attrs, name=dim, _plain=True)
            return out

        def transpose(self, *dims):
            dims = tuple(dims) if dims else self.dims[::-1]
            return self._replace(np.transpose(self._data, [self.dims.index(d) for d in dims]), dims=dims)

        # -- arithmetic with name-based broadcasting
        def _aligned(self, other):
            """Return (a, b, dims, coords): numpy views broadcastable against each other."""
            if not isinstance(other, DataArray):
                return self._data, (other.data if hasattr(other, "data") and not isinstance(other, np.ndarray) else other), self.dims, self._coords
            dims = list(self.dims) + [d for d in other.dims if d not in self.dims]

            def view(da):
                order = [d for d in dims if d in da.dims]
                arr = np.transpose(da._data, [da.dims.index(d) for d in order])
                shape = [da._data.shape[da.dims.index(d)] if d in da.dims else 1 for d in dims]
                return arr.reshape(shape)

            coords = Coordinates()
            for src in (self._coords, other._coords):
                for k, v in src._vars.items():
                    if k not in coords._vars and v.dims != ():
                        coords._vars[k] = v
            return view(self), view(other), tuple(dims), coords

        def _binary_op(self, other, op, reflexive=False):
            if isinstance(other, Dataset):
                return NotImplemented
            a, b, dims, coords = self._aligned(other)
            data = op(b, a) if reflexive else op(a, b)
            out = self._replace(np.asarray(data), dims=dims, coords=coords, keep_attrs=False)
            return out

        __add__ = _binary(operator.add)
        __radd__ = _binary(operator.add, True)
        __sub__ = _binary(operator.sub)
        __rsub__ = _binary(operator.sub, True)
        __mul__ = _binary(operator.mul)
        __rmul__ = _binary(operator.mul, True)
        __truediv__ = _binary(operator.truediv)
        __rtruediv__ = _binary(operator.truediv, True)
        __pow__ = _binary(operator.pow)
        __rpow__ = _binary(operator.pow, True)
        __lt__ = _binary(operator.lt)
        __le__ = _binary(operator.le)
        __gt__ = _binary(operator.gt)
        __ge__ = _binary(operator.ge)
        __eq__ = _binary(operator.eq)
        __ne__ = _binary(operator.ne)
        __and__ = _binary(operator.and_)
        __or__ = _binary(operator.or_)
        __hash__ = None

        def __neg__(self):
            return self._replace(-self._data)

        def __abs__(self):
            return self._replace(np.abs(self._data))

        def __invert__(self):
            return self._replace(~self._data)

        def _inplace(op):  # noqa: N805
            def f(self, other):
                _, b, _, _ = self._aligned(other)
                b = np.asarray(b)
                if b.ndim > self._data.ndim:
                    raise ValueError("in-place operand adds dimensions")
                op(self._data, b)
                return self
            return f

        __iadd__ = _inplace(lambda a, b: np.add(a, b, out=a, casting="unsafe"))
        __isub__ = _inplace(lambda a, b: np.subtract(a, b, out=a, casting="unsafe"))
        __imul__ = _inplace(lambda a, b: np.multiply(a, b, out=a, casting="unsafe"))
        __itruediv__ = _inplace(lambda a, b: np.divide(a, b, out=a, casting="unsafe"))
        del _inplace

        # -- reductions
        def _reduce(self, fn, dim=None, keep_attrs=False, axis=None, **kw):
            if dim is None and axis is not None:
                dim = [self.dims[a] for a in np.atleast_1d(axis)]
            if dim is None:
                axes = None
                dims = ()
            else:
                dl = (dim,) if isinstance(dim, str) else tuple(dim)
                axes = tuple(self.dims.index(d) for d in dl)
                dims = tuple(d for d in self.dims if d not in dl)
            data = np.asarray(fn(self._data, axis=axes, **kw))
            return self._replace(data, dims=dims, keep_attrs=keep_attrs)

        def max(self, dim=None, keep_attrs=False, skipna=True, axis=None, out=None, **_kw):
            fn = _skipna(np.max, np.nanmax) if (skipna and self._data.dtype.kind == "f") else np.max
            return self._reduce(_quiet(fn), dim, keep_attrs, axis=axis)

        def min(self, dim=None, keep_attrs=False, skipna=True, axis=None, out=None, **_kw):
            fn = _skipna(np.min, np.nanmin) if (skipna and self._data.dtype.kind == "f") else np.min
            return self._reduce(_quiet(fn), dim, keep_attrs, axis=axis)

        def mean(self, dim=None, keep_attrs=False, skipna=True, axis=None, out=None, **_kw):
            fn = _skipna(np.mean, np.nanmean) if (skipna and self._data.dtype.kind == "f") else np.mean
            return self._reduce(_quiet(fn), dim, keep_attrs, axis=axis)

        def sum(self, dim=None, keep_attrs=False, skipna=True, axis=None, out=None, **_kw):
            fn = np.nansum if (skipna and self._data.dtype.kind == "f") else np.sum
            return self._reduce(fn, dim, keep_attrs, axis=axis)

        def any(self, dim=None):
            return self._reduce(np.any, dim)

        def all(self, dim=None):
            return self._reduce(np.all, dim)

        def argmax(self, dim=None):
            if dim is None:
                return DataArray(np.asarray(np.nanargmax(self._data)))
            ax = self.dims.index(dim)
            return self._replace(np.nanargmax(self._data, axis=ax), dims=tuple(d for d in self.dims if d != dim), keep_attrs=False)

        # -- masking
        def where(self, cond, other=np.nan, drop=False):
            a, c, dims, coords = self._aligned(cond)
            if isinstance(other, DataArray):
                _, o, _, _ = self._aligned(other)
            else:
                o = other
            c = np.asarray(c, dtype=bool)
            a_b = np.broadcast_to(a, np.broadcast_shapes(np.shape(a), np.shape(c)))
            if o is np.nan or (np.ndim(o) == 0 and isinstance(o, float) and np.isnan(o)):
                base = a_b if a_b.dtype.kind in "fc" else a_b.astype(np.float64)
                data = np.where(c, base, np.nan)
            else:
                data = np.where(c, a_b, o)
            out = self._replace(data, dims=dims, coords=coords)
            if drop:
                cb = np.broadcast_to(c, data.shape)
                for ax, d in enumerate(out.dims):
                    other_axes = tuple(i for i in range(data.ndim) if i != ax)
                    keep = cb.any(axis=other_axes) if other_axes else cb
                    out = out.isel({d: np.flatnonzero(keep)})
                    cb = np.take(cb, np.flatnonzero(keep), axis=ax)
            return out

        def fillna(self, value):
            return self._replace(np.where(np.isnan(self._data), value, self._data))

        def isnull(self):
            return self._replace(np.isnan(self._data), keep_attrs=False)

        # -- interpolation (multi-linear, NaN outside the hull -- xarray's default)
        def interp(self, coords=None, method="linear", **kw):
            req = dict(coords or {}, **kw)
            if method != "linear":
                raise NotImplementedError("only linear interpolation is implemented")
            dims_i = [d for d in self.dims if d in req]
            indexers = {d: req[d] for d in dims_i}
            shared = None
            if all(isinstance(v, DataArray) for v in indexers.values()):
                dsets = {v.dims for v in indexers.values()}
                if len(dsets) == 1 and len(next(iter(dsets))) == 1:
                    shared = next(iter(dsets))[0]
            if shared is not None and shared not in self.dims:
                return self._interp_pointwise(indexers, shared)
            out = self
            for d in dims_i:
                v = indexers[d]
                x_new = np.atleast_1d(np.asarray(v.data if isinstance(v, DataArray) else v, dtype=np.float64))
                out = out._interp_1d(d, x_new, scalar=np.ndim(v) == 0)
            return out

        def _interp_1d(self, dim, x_new, scalar=False):
            ax = self.dims.index(dim)
            x = np.asarray(self._coords[dim].data, dtype=np.float64)
            lo = np.clip(np.searchsorted(x, x_new, side="right") - 1, 0, len(x) - 2)
            w = (x_new - x[lo]) / (x[lo + 1] - x[lo])
            a = np.take(self._data, lo, axis=ax)
            b = np.take(self._data, lo + 1, axis=ax)
            shp = [1] * self.ndim
            shp[ax] = -1
            w_b = w.reshape(shp)
            data = a + (b - a) * w_b
            oob = ((x_new < x[0]) | (x_new > x[-1])).reshape(shp)
            data = np.where(oob, np.nan, data)
            out = self._replace(data)
            out._coords = self._coords.copy(deep=False)
            out._coords._vars[dim] = DataArray(x_new, dims=(dim,), attrs=self._coords[dim].attrs, name=dim, _plain=True)
            if scalar:
                out = out.isel({dim: 0})
            return out

        def _interp_pointwise(self, indexers, new_dim):
            dims_i = list(indexers)
            keep = [d for d in self.dims if d not in dims_i]
            arr = np.transpose(self._data, [self.dims.index(d) for d in dims_i + keep]).astype(np.float64, copy=False)
            npts = indexers[dims_i[0]].size
            lo, w, oob = [], [], np.zeros(npts, dtype=bool)
            for d in dims_i:
                x = np.asarray(self._coords[d].data, dtype=np.float64)
                xn = np.asarray(indexers[d].data, dtype=np.float64)
                l_ = np.clip(np.searchsorted(x, xn, side="right") - 1, 0, len(x) - 2)
                lo.append(l_)
                w.append((xn - x[l_]) / (x[l_ + 1] - x[l_]))
                oob |= (xn < x[0]) | (xn > x[-1])
            res = np.zeros((npts,) + arr.shape[len(dims_i):], dtype=np.float64)
            for corner in range(1 << len(dims_i)):
                wt = np.ones(npts)
                idx = []
                for k in range(len(dims_i)):
                    bit = (corner >> k) & 1
                    idx.append(lo[k] + bit)
                    wt = wt * (w[k] if bit else (1.0 - w[k]))
                vals = arr[tuple(idx)]
                res += vals * wt.reshape((-1,) + (1,) * (vals.ndim - 1))
            res[oob] = np.nan
            out = DataArray.__new__(DataArray)
            out._data = res
            out.dims = (new_dim,) + tuple(keep)
            out.name = self.name
            out.attrs = dict(self.attrs)
            out._coords = Coordinates()
            first = indexers[dims_i[0]]
            if new_dim in first.coords:
                out._coords._vars[new_dim] = first.coords[new_dim]
            for d in dims_i:
                out._coords._vars[d] = DataArray(np.asarray(indexers[d].data), dims=(new_dim,), attrs=self._coords[d].attrs, name=d, _plain=True)
            for d in keep:
                if d in self._coords:
                    out._coords._vars[d] = self._coords[d]
            return out

        def to_dataset(self, name=None):
            return Dataset({name or self.name: self})

    def _skipna(plain, nan_aware):
        """NaN-skipping reduction that only pays for the NaN handling when a NaN is there: the plain reduction
        propagates NaN, so a NaN-free result is already the NaN-skipping one (same summation order for mean)."""
        def g(a, axis=None, **kw):
            r = plain(a, axis=axis, **kw)
            if np.isnan(r).any():
                return nan_aware(a, axis=axis, **kw)
            return r
        return g

    def _quiet(fn):
        def g(a, axis=None, **kw):
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore", RuntimeWarning)
                return fn(a, axis=axis, **kw)
        return g

    class Dataset(Mapping):
        def __init__(self, data_vars: Mapping | None = None, coords=None, attrs=None):
            self._vars: "OrderedDict[str, DataArray]" = OrderedDict()
            self._coords = Coordinates()
            self.attrs = dict(attrs) if attrs else {}
            if coords is not None:
                for k, v in coords.items():
                    self._coords._vars[k] = _coord_var(k, v)
            if data_vars:
                for k, v in data_vars.items():
                    self[k] = v

        def __getitem__(self, key):
            if key in self._vars:
                return self._vars[key]
            if key in self._coords:
                return self._coords[key]
            raise KeyError(key)

        def __setitem__(self, key, value):
            if isinstance(value, tuple):
                dims, data = value[0], np.asarray(value[1])
                attrs = value[2] if len(value) > 2 else None
                dims = (dims,) if isinstance(dims, str) else tuple(dims)
                value = DataArray(data, dims=dims, attrs=attrs,
                                  coords={d: self._coords[d] for d in dims if d in self._coords})
            elif not isinstance(value, DataArray):
                value = DataArray(np.asarray(value))
            for k, c in value.coords._vars.items():
                if k not in self._coords._vars:
                    self._coords._vars[k] = c
            # variables pick up the dataset's index coordinates for their dims
            for d in value.dims:
                if d in self._coords._vars and d not in value.coords._vars:
                    value.coords._vars[d] = self._coords._vars[d]
            out = value.copy(deep=False)
            out.name = key
            self._vars[key] = out

        def __iter__(self):
            return iter(self._vars)

        def __len__(self):
            return len(self._vars)

        def __contains__(self, key):
            return key in self._vars or key in self._coords

        def __repr__(self):
            return f"<Dataset dims={self.sizes} vars={list(self._vars)}>"

        @property
        def data_vars(self):
            return self._vars

        @property
        def coords(self):
            return self._coords

        @property
        def sizes(self):
            out = {}
            for v in self._vars.values():
                for d, n in zip(v.dims, v.shape):
                    out.setdefault(d, n)
            for k, c in self._coords._vars.items():
                if c.dims == (k,):
                    out.setdefault(k, c.size)
            return out

        @property
        def dims(self):
            return self.sizes

        def _map(self, fn):
            out = Dataset(attrs=self.attrs)
            for k, v in self._vars.items():
                out[k] = fn(v)
            for k, c in self._coords._vars.items():
                if k not in out._coords._vars and (c.dims == () or all(d in out.sizes for d in c.dims)):
                    out._coords._vars[k] = c
            return out

        def copy(self, deep=True):
            out = Dataset(attrs=_copy.deepcopy(self.attrs) if deep else dict(self.attrs))
            out._coords = self._coords.copy(deep=deep)
            for k, v in self._vars.items():
                out._vars[k] = v.copy(deep=deep)
            return out

        def __copy__(self):
            return self.copy(deep=False)

        def __deepcopy__(self, memo):
            return self.copy(deep=True)

        def isel(self, indexers=None, **kw):
            idx = dict(indexers or {}, **kw)
            return self._map(lambda v: v.isel({d: k for d, k in idx.items() if d in v.dims}))

        def sel(self, indexers=None, method=None, **kw):
            idx = dict(indexers or {}, **kw)
            return self._map(lambda v: v.sel({d: k for d, k in idx.items() if d in v.dims}, method=method))

        def max(self, dim=None, keep_attrs=False):
            return self._map(lambda v: v.max(dim=dim if dim is None or dim in v.dims else None, keep_attrs=keep_attrs))

        def mean(self, dim=None, keep_attrs=False):
            return self._map(lambda v: v.mean(dim=dim if dim is None or dim in v.dims else None, keep_attrs=keep_attrs))

        def assign_coords(self, coords=None, **kw):
            new = dict(coords or {}, **kw)
            out = self.copy(deep=False)
            for k, v in new.items():
                cv = _coord_var(k, v)
                out._coords._vars[k] = cv
                for name, var in list(out._vars.items()):
                    if all(d in var.dims for d in cv.dims):
                        nv = var.copy(deep=False)
                        nv._coords = var._coords.copy(deep=False)
                        nv._coords._vars[k] = cv
                        out._vars[name] = nv
            return out

        def drop_dims(self, dim):
            dl = (dim,) if isinstance(dim, str) else tuple(dim)
            out = Dataset(attrs=self.attrs)
            for k, c in self._coords._vars.items():
                if not any(d in dl for d in c.dims):
                    out._coords._vars[k] = c
            for k, v in self._vars.items():
                if not any(d in dl for d in v.dims):
                    out._vars[k] = v
            return out

        def drop_vars(self, names):
            nl = (names,) if isinstance(names, str) else tuple(names)
            out = self.copy(deep=False)
            for n in nl:
                out._vars.pop(n, None)
                out._coords._vars.pop(n, None)
            return out

        def load(self):
            return self

        def close(self):
            return None

        def to_netcdf(self, path=None, engine=None, **kw):
            """NetCDF-3 (classic) container through ``scipy.io.netcdf_file`` -- what
            ``engine='scipy'`` writes in xarray.  There is no HDF5 writer in this image, so
            ``engine='h5netcdf'`` (solution.py:515) also lands in the classic format; real
            xarray reads either transparently."""
            from scipy.io import netcdf_file
            f = netcdf_file(path, "w", version=2)
            try:
                for d, n in self.sizes.items():
                    f.createDimension(d, int(n))

                def put(name, da):
                    data = np.asarray(da.data)
                    if data.dtype == np.bool_:
                        data = data.astype(np.int8)
                    if data.dtype == np.int64:
                        data = data.astype(np.int32) if np.all(np.abs(data) < 2 ** 31) else data.astype(np.float64)
                    v = f.createVariable(name, data.dtype, tuple(da.dims))
                    v[...] = data
                    for k, a in da.attrs.items():
                        if isinstance(a, (str, int, float, np.integer, np.floating)):
                            setattr(v, k, a)
                for k, c in self._coords._vars.items():
                    put(k, c)
                for k, v in self._vars.items():
                    put(k, v)
                for k, a in self.attrs.items():
                    if isinstance(a, (str, int, float)):
                        setattr(f, k, a)
            finally:
                f.close()

    def concat(objs: Iterable, dim: str):
        objs = list(objs)
        if not objs:
            raise ValueError("must supply at least one object to concatenate")
        if isinstance(objs[0], Dataset):
            out = Dataset(attrs=objs[0].attrs)
            for k in objs[0]._vars:
                out[k] = concat([o[k] for o in objs], dim)
            return out
        pieces = []
        labels = []
        for o in objs:
            if dim in o.dims:
                pieces.append(o)
                labels.extend(np.asarray(o.coords[dim].data).tolist() if dim in o.coords else range(o.sizes[dim]))
            else:
                pieces.append(o.expand_dims(dim, 0))
                c = o.coords._vars.get(dim)
                labels.append(c.data.item() if c is not None else len(labels))
        first = pieces[0]
        ax = first.dims.index(dim)
        arrs = [np.asarray(p.data) for p in pieces]
        if (ax == 0 and first.ndim > 2 and all(a.shape[0] == 1 and a.dtype == arrs[0].dtype and a.shape == arrs[0].shape
                                                and a[0].flags.f_contiguous and not a[0].flags.c_contiguous for a in arrs)):
            # Fortran-ordered pieces (what run_simulation returns, kwave_if.py:132-141 reshape(order='F')): keep every
            # piece's memory order inside the stack, so stacking is one block copy per piece instead of a strided transpose
            buf = np.empty((len(arrs),) + arrs[0].shape[1:][::-1], dtype=arrs[0].dtype)
            data = buf.transpose((0,) + tuple(range(first.ndim - 1, 0, -1)))
            for i, a in enumerate(arrs):
                data[i] = a[0]
        else:
            data = np.concatenate(arrs, axis=ax)
        out = first._replace(data)
        out._coords = Coordinates()
        for k, v in first._coords._vars.items():
            if k != dim and dim not in v.dims:
                out._coords._vars[k] = v
        out._coords._vars[dim] = DataArray(np.asarray(labels), dims=(dim,), name=dim, _plain=True)
        return out

    def open_dataset(filename_or_obj, engine=None, **k):
        """Read a NetCDF-3 file (path or bytes) written by ``Dataset.to_netcdf``."""
        import io
        from scipy.io import netcdf_file
        src = io.BytesIO(filename_or_obj) if isinstance(filename_or_obj, (bytes, bytearray)) else str(filename_or_obj)
        f = netcdf_file(src, "r", mmap=False)
        try:
            def attrs_of(v):
                out = {}
                for k2, a in v._attributes.items():
                    out[k2] = a.decode() if isinstance(a, bytes) else (a.item() if isinstance(a, np.ndarray) and a.size == 1 else a)
                return out
            coords, data = {}, {}
            for name, v in f.variables.items():
                arr = np.array(v[...])
                if arr.dtype.byteorder == ">":
                    arr = arr.astype(arr.dtype.newbyteorder("="))
                da = DataArray(arr, dims=tuple(v.dimensions), name=name, attrs=attrs_of(v), _plain=True)
                (coords if (len(v.dimensions) == 1 and v.dimensions[0] == name) else data)[name] = da
            ds = Dataset(coords=coords)
            for name, da in data.items():
                ds[name] = DataArray(da.data, coords={d: coords[d] for d in da.dims if d in coords}, dims=da.dims, name=name,
                                     attrs=da.attrs)
            return ds
        finally:
            f.close()

"""Transducer: a list of elements plus per-element drive synthesis.

API mirror of /root/reference/src/openlifu/xdc/transducer.py (``Transducer:19``,
``calc_output:95-112``, ``get_effective_origin:191``, ``get_positions:204``,
``convert_transform:210``, ``merge:231``, ``gen_matrix_array:372-406``,
``TransformedTransducer:409``) without the vtk drawing helpers (out of the hot path).
"""
from __future__ import annotations

import copy
import logging
import json
from dataclasses import dataclass, field
from typing import Any, Dict, List

import numpy as np

from ..util.units import getunitconversion
from .element import Element

DIMS = ["x", "y", "z"]


def _axis_rotation(dim: str, angle_rad: float) -> np.ndarray:
    c, s = np.cos(angle_rad), np.sin(angle_rad)
    m = np.eye(4)
    i, j = {"x": (1, 2), "y": (2, 0), "z": (0, 1)}[dim]
    m[i, i] = c
    m[j, j] = c
    m[i, j] = -s
    m[j, i] = s
    return m


_WARNED_ELEMENT_SENSITIVITY = False


def _warn_element_sensitivity():
    global _WARNED_ELEMENT_SENSITIVITY
    if not _WARNED_ELEMENT_SENSITIVITY:
        _WARNED_ELEMENT_SENSITIVITY = True
        logging.getLogger(__name__).warning(
            "Transducer has per-element sensitivities: openlifu_b200 applies each element's own sensitivity, whereas the "
            "reference's in-place multiplication compounds them element after element (xdc/element.py:145-153); source "
            "amplitudes differ from the reference for such arrays")


@dataclass
class Transducer:
    id: str = "transducer"
    name: str = ""
    elements: List[Element] = field(default_factory=list)
    frequency: float = 400.6e3
    units: str = "m"
    attrs: Dict[str, Any] = field(default_factory=dict)
    registration_surface_filename: str | None = None
    transducer_body_filename: str | None = None
    standoff_transform: np.ndarray = field(default_factory=lambda: np.eye(4, dtype=float))
    sensitivity: float | None = None
    impulse_response: np.ndarray | None = None
    impulse_dt: float | None = None
    module_invert: List[bool] = field(default_factory=lambda: [False])

    def __post_init__(self):
        if self.name == "":
            self.name = self.id
        for el in self.elements:
            el.rescale(self.units)
        if self.impulse_response is not None:
            self.impulse_response = np.array(self.impulse_response, dtype=np.float64)
            if self.impulse_response.ndim != 1 or len(self.impulse_response) < 2:
                raise ValueError("Impulse response must be a 1-dimensional array.")
            if self.impulse_dt is None:
                raise ValueError("Impulse response timestep must be set if impulse response is set.")

    # ---------------------------------------------------------------- drive synthesis
    def drive_plan(self, dt, delays=None, apod=None):
        """What the solver needs instead of the dense (n_el, L) matrix: integer delay samples
        ``int(delay/dt)`` (truncation, transducer.py:107) and per-element gains
        ``apod * element gain`` ; the transducer sensitivity scales the base signal."""
        n = self.numelements()
        delays = np.zeros(n) if delays is None else np.asarray(delays, dtype=np.float64)
        apod = np.ones(n) if apod is None else np.asarray(apod, dtype=np.float64)
        if self.impulse_response is not None:
            raise NotImplementedError("array impulse responses are not supported on the simulation path")
        n_delay = np.array([int(dl / dt) for dl in delays], dtype=np.int32)
        gains = np.array([a * el.scalar_gain() for a, el in zip(apod, self.elements)], dtype=np.float64)
        base_gain = 1.0 if self.sensitivity is None else float(self.sensitivity)
        if any(el.sensitivity is not None for el in self.elements):
            # documented deviation (INTEGRATION.md "Per-element sensitivity"): the reference multiplies the SHARED
            # input array in place (element.py:145-153), so element k is scaled by the product of the sensitivities of
            # elements 0..k; here every element gets its own sensitivity only
            _warn_element_sensitivity()
        return n_delay, gains, base_gain

    def calc_output(self, input_signal, dt, delays: np.ndarray = None, apod: np.ndarray = None):
        """(n_elements, L) drive matrix: element e emits gain_e * signal delayed by int(delay_e/dt)."""
        sig = np.asarray(input_signal, dtype=np.float64)
        n_delay, gains, base_gain = self.drive_plan(dt, delays, apod)
        out = np.zeros((self.numelements(), int(n_delay.max(initial=0)) + sig.size))
        for e, (nd, g) in enumerate(zip(n_delay, gains)):
            out[e, nd:nd + sig.size] = g * (sig * base_gain)
        return out

    # ---------------------------------------------------------------- geometry
    def copy(self):
        return copy.deepcopy(self)

    def numelements(self):
        return len(self.elements)

    def get_area(self, units=None):
        units = self.units if units is None else units
        return sum(el.get_area(units=units) for el in self.elements)

    def get_corners(self, transform=None, units=None):
        units = self.units if units is None else units
        m = np.eye(4) if transform is None else transform
        return [el.get_corners(units=units, matrix=m) for el in self.elements]

    def get_positions(self, transform=None, units=None):
        units = self.units if units is None else units
        m = np.eye(4) if transform is None else transform
        return np.array([el.get_position(units=units, matrix=m) for el in self.elements])

    def get_effective_origin(self, apodizations: np.ndarray, units: str | None = None):
        """Apodization-weighted centroid of the element positions."""
        w = np.asarray(apodizations, dtype=np.float64).reshape(-1, 1)
        return (w * self.get_positions(units=units)).sum(axis=0) / w.sum()

    def convert_transform(self, matrix: np.ndarray, units: str) -> np.ndarray:
        out = matrix.copy()
        out[0:3, 3] *= getunitconversion(units, self.units)
        return out

    def get_standoff_transform_in_units(self, units: str) -> np.ndarray:
        out = self.standoff_transform.copy()
        out[0:3, 3] *= getunitconversion(self.units, units)
        return out

    def rescale(self, units):
        if self.units != units:
            for el in self.elements:
                el.rescale(units)
            self.units = units

    def sort_by_index(self):
        self.elements = [self.elements[i] for i in np.argsort([el.index for el in self.elements])]

    def sort_by_pin(self):
        self.elements = [self.elements[i] for i in np.argsort([el.pin for el in self.elements])]

    def transform(self, matrix, units=None):
        if units is not None:
            self.rescale(units)
        inv = np.linalg.inv(matrix)
        for el in self.elements:
            el.set_matrix(inv @ el.get_matrix())

    def translate(self, dim, amount: float, units=None):
        m = np.eye(4)
        m[DIMS.index(dim), 3] = amount
        self.transform(m, units=units)

    def rotate(self, dim, angle: float, units="deg"):
        self.transform(_axis_rotation(dim, np.deg2rad(angle) if units == "deg" else angle))

    @staticmethod
    def merge(list_of_transducers, offset_pins=False, offset_indices=False,
              merge_mismatched_sensitivity=True, merged_attrs: dict | None = None) -> "Transducer":
        arrays = [a.copy() for a in list_of_transducers]
        sens = np.array([a.sensitivity for a in arrays if a.sensitivity is not None])
        if 0 < len(sens) < len(arrays):
            raise ValueError("If one transducer has a sensitivity, all must have a sensitivity.")
        if len(set(sens)) > 1:
            if not merge_mismatched_sensitivity:
                raise ValueError("Transducers have different sensitivities. Use merge_mismatched_sensitivity=True "
                                 "to merge the relative sensitivities into the merged elements")
            top = sens.max()
            for a, rel in zip(arrays, sens / top):
                for el in a.elements:
                    el.sensitivity = rel if el.sensitivity is None else el.sensitivity * rel
                a.sensitivity = top
        merged = arrays[0]
        for other in arrays[1:]:
            n0 = merged.numelements()
            for el in other.elements:
                if offset_pins:
                    el.pin += n0
                if offset_indices:
                    el.index += n0
            merged.elements += other.elements
            merged.module_invert += other.module_invert
        for k, v in (merged_attrs or {}).items():
            setattr(merged, k, v)
        return merged

    # ---------------------------------------------------------------- (de)serialisation
    def to_dict(self):
        d = dict(self.__dict__)
        d["elements"] = [el.to_dict() for el in self.elements]
        if self.impulse_response is None:
            d.pop("impulse_response")
        else:
            d["impulse_response"] = self.impulse_response.tolist()
        if self.impulse_dt is None:
            d.pop("impulse_dt")
        d["standoff_transform"] = np.asarray(self.standoff_transform).tolist()
        return d

    @staticmethod
    def from_dict(d, **kwargs):
        d = dict(d)
        d["elements"] = [Element.from_dict(e) for e in d["elements"]]
        ir = d.get("impulse_response")
        if ir is not None:
            if len(ir) == 1 and "sensitivity" not in d:
                d["sensitivity"] = ir[0]
                del d["impulse_response"]
            else:
                d["impulse_response"] = np.array(ir)
        if d.get("standoff_transform") is not None:
            d["standoff_transform"] = np.array(d["standoff_transform"])
        return Transducer(**d, **kwargs)

    @staticmethod
    def from_file(filename):
        with open(filename) as f:
            return Transducer.from_dict(json.load(f))

    @staticmethod
    def from_json(json_string: str) -> "Transducer":
        return Transducer.from_dict(json.loads(json_string))

    def to_json(self, compact: bool = False) -> str:
        return json.dumps(self.to_dict(), separators=(",", ":")) if compact else json.dumps(self.to_dict(), indent=4)

    def to_file(self, filename):
        with open(filename, "w") as f:
            f.write(self.to_json())

    @staticmethod
    def gen_matrix_array(nx=2, ny=2, pitch=1, kerf=0, units="mm", **kwargs):
        """Flat nx x ny matrix array centred on the origin; element i sits in column i // ny,
        row i % ny (rows run from +y to -y); indices and pins are 1-based."""
        cols = (np.arange(nx) - (nx - 1) / 2) * pitch
        rows = -(np.arange(ny) - (ny - 1) / 2) * pitch
        side = pitch - kerf
        elements = [Element(index=i + 1, pin=i + 1, position=np.array([cols[i // ny], rows[i % ny], 0]),
                            orientation=np.zeros(3), size=np.array([side, side]), units=units)
                    for i in range(nx * ny)]
        return Transducer(elements=elements, units=units, **kwargs)


@dataclass
class TransformedTransducer(Transducer):
    transform: np.ndarray = field(default_factory=lambda: np.eye(4))

    def bake(self):
        d = self.to_dict()
        d.pop("transform")
        t = Transducer.from_dict(d)
        t.transform(self.transform, units=self.units)
        return t

    def _shift(self, dim, amount):
        m = np.eye(4)
        m[DIMS.index(dim), 3] = amount
        return np.linalg.inv(m)

    def translate_global(self, dim, amount, units=None):
        self.transform = self.transform @ self._shift(dim, amount)

    def translate_local(self, dim, amount, units=None):
        self.transform = self._shift(dim, amount) @ self.transform

    def rotate_global(self, dim, angle: float, units="deg"):
        self.transform = self.transform @ _axis_rotation(dim, np.deg2rad(angle) if units == "deg" else angle)

    def rotate_local(self, dim, angle: float, units="deg"):
        self.transform = _axis_rotation(dim, np.deg2rad(angle) if units == "deg" else angle) @ self.transform

"""Transducer element geometry and per-element drive output.

API mirror of /root/reference/src/openlifu/xdc/element.py (``Element:33``, ``calc_output:144``,
``get_position:166``, ``get_size:174``, ``get_matrix:200``, ``get_angle:216``,
``distance_to_point:239``, ``angle_to_point:248``).  Orientation is (az, el, roll) in radians
about the (y, x', z'') axes; the element frame is Ry(az) Rx(el) Rz(roll).
"""
from __future__ import annotations

import copy
from dataclasses import dataclass, field

import numpy as np

from ..util.units import getunitconversion


def _rot_y(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, 0, s], [0, 1, 0], [-s, 0, c]])


def _rot_x(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[1, 0, 0], [0, c, -s], [0, s, c]])


def _rot_z(a):
    c, s = np.cos(a), np.sin(a)
    return np.array([[c, -s, 0], [s, c, 0], [0, 0, 1]])


def matrix2xyz(matrix):
    """Inverse of ``Element.get_matrix``: (x, y, z, az, el, roll) of a 4x4 frame (element.py:13-31)."""
    x, y, z = matrix[0, 3], matrix[1, 3], matrix[2, 3]
    az = np.arctan2(matrix[0, 2], matrix[2, 2])
    el = -np.arctan2(matrix[1, 2], np.hypot(matrix[2, 2], matrix[0, 2]))
    r = _rot_y(az) @ _rot_x(el)
    xv = matrix[:3, 0]
    roll = np.arctan2(xv @ r[:3, 1], xv @ r[:3, 0])
    return x, y, z, az, el, roll


@dataclass
class Element:
    index: int = 0
    position: np.ndarray = field(default_factory=lambda: np.array([0., 0., 0.]))
    orientation: np.ndarray = field(repr=False, default_factory=lambda: np.array([0., 0., 0.]))
    size: np.ndarray = field(default_factory=lambda: np.array([1., 1.]))
    sensitivity: float | None = None
    impulse_response: np.ndarray | None = None
    impulse_dt: float | None = None
    pin: int = -1
    units: str = "mm"

    def __post_init__(self):
        self.position = np.array(self.position, dtype=np.float64)
        self.orientation = np.array(self.orientation, dtype=np.float64)
        self.size = np.array(self.size, dtype=np.float64)
        if self.position.shape != (3,):
            raise ValueError("Position must be a 3-element array.")
        if self.orientation.shape != (3,):
            raise ValueError("Orientation must be a 3-element array.")
        if self.size.shape != (2,):
            raise ValueError("Size must be a 2-element array.")
        if self.impulse_response is not None:
            ir = self.impulse_response
            if isinstance(ir, (int, float)):
                ir = [ir]
            self.impulse_response = np.array(ir, dtype=np.float64)
            if self.impulse_response.ndim != 1:
                raise ValueError("Impulse response must be a 1-dimensional array.")
            if len(self.impulse_response) > 1 and self.impulse_dt is None:
                raise ValueError("Impulse response timestep must be set if impulse response is an array.")

    # named access to the packed vectors ------------------------------------------------
    def _vec_property(vec, i):  # noqa: N805
        def get(self):
            return getattr(self, vec)[i]

        def set_(self, value):
            getattr(self, vec)[i] = value
        return property(get, set_)

    x = _vec_property("position", 0)
    y = _vec_property("position", 1)
    z = _vec_property("position", 2)
    az = _vec_property("orientation", 0)
    el = _vec_property("orientation", 1)
    roll = _vec_property("orientation", 2)
    width = _vec_property("size", 0)
    length = _vec_property("size", 1)
    del _vec_property

    # drive ----------------------------------------------------------------------------
    def scalar_gain(self) -> float:
        """Product of the scalar impulse response and the element sensitivity: what
        ``calc_output`` multiplies the drive signal by.  Array impulse responses are not
        usable in the reference either."""
        g = 1.0
        if self.impulse_response is not None:
            if len(self.impulse_response) != 1:
                raise NotImplementedError("array impulse responses are not supported on the simulation path")
            g *= float(self.impulse_response[0])
        if self.sensitivity is not None:
            g *= float(self.sensitivity)
        return g

    def calc_output(self, input_signal, dt):
        return np.asarray(input_signal) * self.scalar_gain()

    # geometry -------------------------------------------------------------------------
    def copy(self):
        return copy.deepcopy(self)

    def rescale(self, units):
        if self.units != units:
            k = getunitconversion(self.units, units)
            self.position *= k
            self.size *= k
            self.units = units

    def _scale_to(self, units):
        return getunitconversion(self.units, self.units if units is None else units)

    def get_position(self, units=None, matrix=np.eye(4)):
        return (matrix @ np.append(self.position * self._scale_to(units), 1.0))[:3]

    def get_size(self, units=None):
        k = self._scale_to(units)
        return self.size[0] * k, self.size[1] * k

    def get_area(self, units=None):
        w, l = self.get_size(units)
        return w * l

    def get_matrix(self, units=None):
        m = np.eye(4)
        m[:3, :3] = _rot_y(self.az) @ (_rot_x(self.el) @ _rot_z(self.roll))
        m[:3, 3] = self.get_position(units=units)
        return m

    def get_corners(self, units=None, matrix=np.eye(4)):
        k = self._scale_to(units)
        hw, hl = 0.5 * self.width, 0.5 * self.length
        local = np.array([[-hw, -hw, hw, hw], [-hl, hl, hl, -hl], [0, 0, 0, 0], [1, 1, 1, 1]], dtype=np.float64)
        return (matrix @ (self.get_matrix() @ local))[:3] * k

    def get_angle(self, units="rad"):
        """(el, az, roll): rotations about x, y', z''."""
        ang = np.array([self.el, self.az, self.roll])
        if units == "deg":
            ang = np.degrees(ang)
        return ang[0], ang[1], ang[2]

    def distance_to_point(self, point, units=None, matrix=np.eye(4)):
        g = matrix @ np.append(self.get_position(units=units), 1.0)
        return np.linalg.norm(np.asarray(point) - g[:3], 2)

    def angle_to_point(self, point, units=None, return_as="rad", matrix=np.eye(4)):
        frame = matrix @ self.get_matrix(units=units)
        ray = np.asarray(point) - frame[:3, 3]
        normal = frame[:3, 2]
        ray = ray / np.linalg.norm(ray, 2)
        normal = normal / np.linalg.norm(normal, 2)
        theta = np.arcsin(np.linalg.norm(np.cross(ray, normal), 2))
        return np.degrees(theta) if return_as == "deg" else theta

    def set_matrix(self, matrix, units=None):
        if units is not None:
            self.rescale(units)
        x, y, z, az, el, roll = matrix2xyz(matrix)
        self.position = np.array([x, y, z])
        self.orientation = np.array([az, el, roll])

    # (de)serialisation ----------------------------------------------------------------
    def to_dict(self):
        d = {"index": self.index, "position": self.position.tolist(), "orientation": self.orientation.tolist(),
             "size": self.size.tolist(), "pin": self.pin, "units": self.units}
        if self.impulse_response is not None:
            d["impulse_response"] = self.impulse_response.tolist()
        if self.impulse_dt is not None:
            d["impulse_dt"] = self.impulse_dt
        return d

    @staticmethod
    def from_dict(d):
        d = copy.deepcopy(d)
        if "x" in d:   # legacy flat layout
            d["position"] = np.array([d.pop("x"), d.pop("y"), d.pop("z")])
            d["orientation"] = np.array([d.pop("az"), d.pop("el"), d.pop("roll")])
            d["size"] = np.array([d.pop("w"), d.pop("l")])
        if d.get("impulse_response") is not None:
            d["impulse_response"] = np.array(d["impulse_response"])
        if d.get("impulse_dt") is not None:
            d["impulse_dt"] = float(d["impulse_dt"])
        return Element(**d)

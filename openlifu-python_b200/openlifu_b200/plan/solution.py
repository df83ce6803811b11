"""``Solution``: beamforming result + stacked simulation output + analysis / scaling.

Mirrors /root/reference/src/openlifu/plan/solution.py (``Solution:38``, ``analyze:135-281``,
``compute_scaling_factors:283-311``, ``scale:313-338``, duty cycles ``:340-363``, ``get_ita:365-388``,
(de)serialisation ``:390-533``): same fields, method names, return values and in-place semantics
(``scale`` rescales ``simulation_result`` through ``.data`` so the arrays must be writable numpy).

Documented deviations from reference bugs: ``analyze`` does not let
``Transducer.calc_output`` compound the transducer sensitivity into the shared input signal once per
focus, and ``get_ita`` does not build the ``(F, x, y, z, F)`` temporary -- its result (each focus'
intensity times the two duty cycles) is the same.
"""
from __future__ import annotations

import base64
import json
import logging
import os
import tempfile
from dataclasses import asdict, dataclass, field, replace
from datetime import datetime
from pathlib import Path
from typing import Dict, List, Tuple

import numpy as np

from .. import xa
from ..bf import Pulse, Sequence
from ..bf.focal_patterns import FocalPattern
from ..geo import Point
from ..util.units import getunitconversion, rescale_coords, rescale_data_arr
from ..xdc import Transducer
from .param_constraint import ParameterConstraint
from .solution_analysis import (
    FocusFrame,
    SolutionAnalysis,
    SolutionAnalysisOptions,
    _axes,
    _bounds_from_line,
    trilinear_line,
)


class _Encoder(json.JSONEncoder):
    """numpy / datetime / dataclass aware encoder (the role of util/json.py:PYFUSEncoder)."""

    def default(self, o):
        if isinstance(o, np.integer):
            return int(o)
        if isinstance(o, np.floating):
            return float(o)
        if isinstance(o, np.ndarray):
            return o.tolist()
        if isinstance(o, datetime):
            return o.isoformat()
        if hasattr(o, "to_dict"):
            return o.to_dict()
        return super().default(o)


def _nc_path_for(json_filepath: Path) -> Path:
    return json_filepath.parent / (json_filepath.name.split(".")[0] + ".nc")


def _ratio(side: float, main: float) -> float:
    if main == 0:
        return np.inf if side != 0 else np.nan
    return side / main


def _masked_max(values: np.ndarray, mask: np.ndarray) -> float:
    """max over ``mask`` ignoring NaN; NaN when the selection is empty (``where().max()`` semantics)."""
    sel = values[..., mask] if values.ndim > mask.ndim else values[mask]
    sel = sel[~np.isnan(sel)]
    return float(sel.max()) if sel.size else float("nan")


def _pick_engine(engine, simulation_result) -> str:
    """``engine`` / ``$LIFU_ANALYZE`` when given; otherwise the device engine when a GPU and the built library are there
    and the fields have the layout the kernels take, else the host (numpy) evaluation -- announced at WARNING level, it
    is the reference's own algorithm but 30x slower on a 216^3 grid."""
    engine = engine or os.environ.get("LIFU_ANALYZE") or None
    if engine is None:
        from ..util.checkgpu import gpu_available
        why = None
        try:
            if np.asarray(simulation_result["p_min"].data).dtype != np.float32 or simulation_result["p_min"].ndim != 4:
                why = "the fields are not a float32 (focus, x, y, z) stack"
            elif not gpu_available():
                why = "no CUDA device is visible"
            else:
                from .. import _lib
                _lib.load()
        except Exception as e:  # noqa: BLE001
            why = f"the device engine is unavailable ({e})"
        engine = "cuda" if why is None else "host"
        if why is not None:
            logging.getLogger(__name__).warning("Solution.analyze: using the host (numpy) engine because %s; pass "
                                                "engine='host' to select it explicitly", why)
    if engine not in ("cuda", "host"):
        raise ValueError(f"Unknown analysis engine '{engine}' (expected 'cuda' or 'host')")
    return engine


@dataclass
class Solution:
    id: str = "solution"
    name: str = "Solution"
    protocol_id: str | None = None
    transducer: Transducer | None = None
    date_created: datetime = field(default_factory=datetime.now)
    description: str = ""
    delays: np.ndarray | None = None
    apodizations: np.ndarray | None = None
    pulse: Pulse = field(default_factory=Pulse)
    voltage: float = 1.0
    sequence: Sequence = field(default_factory=Sequence)
    foci: List[Point] = field(default_factory=list)
    target: Point | None = None
    simulation_result: "xa.Dataset" = field(default_factory=xa.Dataset)
    approved: bool = False

    def __post_init__(self):
        if self.delays is not None:
            self.delays = np.array(self.delays, ndmin=2)
        if self.apodizations is not None:
            self.apodizations = np.array(self.apodizations, ndmin=2)
        if self.pulse.frequency <= 0:
            raise ValueError("Pulse frequency must be positive")
        if self.voltage <= 0:
            raise ValueError("Voltage must be positive")
        seq = self.sequence
        if seq.pulse_interval <= 0:
            raise ValueError("Pulse interval must be positive")
        if seq.pulse_count <= 0:
            raise ValueError("Pulse count must be positive")
        if seq.pulse_train_interval < 0:
            raise ValueError("Pulse train interval must be non-negative")
        if 0 < seq.pulse_train_interval < seq.pulse_interval * seq.pulse_count:
            raise ValueError("Pulse train interval must be greater than or equal to the total pulse interval")
        if seq.pulse_train_count <= 0:
            raise ValueError("Pulse train count must be positive")
        nf = len(self.foci)
        if nf > 0 and self.delays is not None and self.delays.shape[0] != nf:
            raise ValueError(f"Delays number of foci ({self.delays.shape[0]}) does not match number of foci ({nf})")
        if nf > 0 and self.apodizations is not None and self.apodizations.shape[0] != nf:
            raise ValueError(f"Apodizations number of foci ({self.apodizations.shape[0]}) does not match number of foci ({nf})")
        if self.delays is not None and self.apodizations is not None:
            if self.apodizations.shape[0] != self.delays.shape[0]:
                raise ValueError(f"Apodizations number of foci ({self.apodizations.shape[0]}) does not match delays "
                                 f"number of foci ({self.delays.shape[0]})")
            if self.apodizations.shape[1] != self.delays.shape[1]:
                raise ValueError(f"Apodizations number of elements {self.apodizations.shape[1]} does not match delays "
                                 f"shape ({self.delays.shape[1]})")

    def num_foci(self) -> int:
        return len(self.foci)

    # ------------------------------------------------------------------------------ analysis
    def analyze(self, options: SolutionAnalysisOptions = SolutionAnalysisOptions(),
                param_constraints: Dict[str, ParameterConstraint] | None = None,
                engine: str | None = None) -> SolutionAnalysis:
        """Beam metrics per focus (reference ``analyze``, solution.py:135-281).

        One pass per focus over the grid: the focus-frame distance map gives the main-lobe /
        side-lobe selections, six trilinear line scans give the beam widths.

        ``engine`` (not in the reference): ``"cuda"`` evaluates the O(V) passes with the ``lifu_analysis_*``
        kernels of liblifusim and raises if the library or a GPU is missing; ``"host"`` is the
        numpy evaluation the reference's own analysis corresponds to.  ``None``: ``$LIFU_ANALYZE`` if set, else
        "cuda" when a GPU is visible and the fields have the solver's dtypes, else "host"."""
        if getattr(self, "_stack", None) is not None:
            engine = "cuda"        # the fields live on the device (Protocol.calc_solution(on_device=True))
        engine = _pick_engine(engine, self.simulation_result)
        if engine == "cuda":
            # called through the class so that `analyze` can be grafted onto the reference's own Solution by name
            return Solution._analyze_cuda(self, options, param_constraints)
        out = SolutionAnalysis()
        units = options.distance_units
        dt = 1 / (self.pulse.frequency * 20)
        input_signal_V = self.pulse.calc_pulse(self.pulse.calc_time(dt)) * self.voltage

        pnp_all = rescale_data_arr(rescale_coords(self.simulation_result["p_min"], units), "MPa")
        ipa_all = rescale_data_arr(rescale_coords(self.simulation_result["intensity"], units), "W/cm^2")
        ita_all = rescale_coords(self.get_ita(units="mW/cm^2"), units)
        if options.sidelobe_radius is np.nan:
            options.sidelobe_radius = options.mainlobe_radius

        standoff_Z = options.standoff_density * 1500
        c_tic = 40e-3   # W/cm
        d_eq_cm = np.sqrt(4 * self.transducer.get_area("cm") / np.pi)
        ele_sizes_cm2 = np.array([el.get_area("cm") for el in self.transducer.elements])

        out.duty_cycle_pulse_train_pct = self.get_pulsetrain_dutycycle() * 100
        out.duty_cycle_sequence_pct = self.get_sequence_dutycycle() * 100
        seq = self.sequence
        if seq.pulse_train_interval == 0:
            out.sequence_duration_s = float(seq.pulse_interval * seq.pulse_count * seq.pulse_train_count)
        else:
            out.sequence_duration_s = float(seq.pulse_train_interval * seq.pulse_train_count)

        pnp_native = np.asarray(pnp_all.data)          # float32 when it comes from the solver
        pnp_v = pnp_native.astype(np.float64, copy=False)
        ipa_v = np.asarray(ipa_all.data, dtype=np.float64)
        ita_v = np.asarray(ita_all.data, dtype=np.float64)
        space = pnp_all.isel(focal_point_index=0)
        axes = _axes(space)
        dims = list(space.dims)
        to_mm_axes = [getunitconversion(space.coords[d].attrs.get("units", units), "mm") for d in dims]
        z_ok = axes[2] > options.sidelobe_zmin
        z_sel = np.broadcast_to(z_ok[None, None, :], pnp_v.shape[1:])
        ar = options.mainlobe_aspect_ratio

        power_W = np.zeros(self.num_foci())
        TIC = np.zeros(self.num_foci())
        for i in range(self.num_foci()):
            pnp, ipa = pnp_v[i], ipa_v[i]
            focus = self.foci[i].get_position(units=units)
            focus_mm = self.foci[i].get_position(units="mm")
            out.target_position_lat_mm += [focus_mm[0]]
            out.target_position_ele_mm += [focus_mm[1]]
            out.target_position_ax_mm += [focus_mm[2]]
            apod = self.apodizations[i]
            origin = self.transducer.get_effective_origin(apodizations=apod, units=units)
            p0_Pa = np.max(self.transducer.calc_output(input_signal_V.copy(), dt, delays=self.delays[i, :], apod=apod), axis=1)

            frame = FocusFrame(focus, origin)
            dist = frame.distance(axes, ar)
            main = dist < options.mainlobe_radius
            side = (dist > options.sidelobe_radius) & z_sel

            pk = _masked_max(pnp, main)
            # -3 dB centroid of the main lobe, in mm
            # (weights and their total keep the field's own dtype, as the reference's DataArray sum does)
            with np.errstate(invalid="ignore"):
                w = np.where(main & (pnp_native[i] > pk * 10 ** (-3 / 20)), pnp_native[i], pnp_native.dtype.type(0))
            tot = w.sum()
            with np.errstate(invalid="ignore", divide="ignore"):
                cen = [float(np.sum(w * a.reshape([-1 if k == j else 1 for j in range(3)])) / tot * s)
                       for k, (a, s) in enumerate(zip(axes, to_mm_axes))]
            out.focal_centroid_lat_mm += [cen[0]]
            out.focal_centroid_ele_mm += [cen[1]]
            out.focal_centroid_ax_mm += [cen[2]]
            out.mainlobe_pnp_MPa += [pk]

            to_mm = getunitconversion(units, "mm")
            for k, (named, scale) in enumerate(zip(("lat", "ele", "ax"), ar)):
                n = pnp.shape[k] * 2
                offs = np.linspace(-scale * options.beamwidth_radius, scale * options.beamwidth_radius, n)
                line = trilinear_line(pnp, axes, frame.line(k, offs))
                for db in (3, 6):
                    neg, pos = _bounds_from_line(offs, line, pk * 10 ** (-db / 20))
                    name = f"beamwidth_{named}_{db}dB_mm"
                    setattr(out, name, [*getattr(out, name), to_mm * (pos - neg)])

            out.mainlobe_isppa_Wcm2 += [_masked_max(ipa, main)]
            out.mainlobe_ispta_mWcm2 += [_masked_max(ita_v, main)]      # over every focus' field, as the reference does
            side_pnp, side_ipa = _masked_max(pnp, side), _masked_max(ipa, side)
            out.sidelobe_pnp_MPa += [side_pnp]
            out.sidelobe_isppa_Wcm2 += [side_ipa]
            out.sidelobe_to_mainlobe_pressure_ratio += [_ratio(side_pnp, out.mainlobe_pnp_MPa[-1])]
            out.sidelobe_to_mainlobe_intensity_ratio += [_ratio(side_ipa, out.mainlobe_isppa_Wcm2[-1])]
            out.global_pnp_MPa += [_masked_max(pnp, z_sel)]
            out.global_isppa_Wcm2 += [_masked_max(ipa, z_sel)]

            i0ta_Wcm2 = (p0_Pa ** 2 / (2 * standoff_Z)) * 1e-4 * out.duty_cycle_sequence_pct / 100
            power_W[i] = np.mean(np.sum(i0ta_Wcm2 * ele_sizes_cm2 * self.apodizations[i, :]))
            TIC[i] = power_W[i] / (d_eq_cm * c_tic)
            out.p0_MPa += [1e-6 * float(np.max(p0_Pa))]

        out.global_ispta_mWcm2 = float(np.nanmax(ita_v * z_ok[None, None, None, :]))
        out.MI = float(np.max(out.mainlobe_pnp_MPa) / np.sqrt(self.pulse.frequency * 1e-6))
        out.TIC = float(np.mean(TIC))
        out.voltage_V = self.voltage
        out.power_W = float(np.mean(power_W))
        out.param_constraints = {} if param_constraints is None else param_constraints
        return out

    def _analyze_cuda(self, options: SolutionAnalysisOptions,
                      param_constraints: Dict[str, ParameterConstraint] | None = None) -> SolutionAnalysis:
        """``analyze`` with every pass over the fields on the GPU (``lifu_analysis_*``, csrc/analysis.cu).

        The fields are staged once in their stored units (float32 Pa, float64 W/cm^2): the unit factors are
        monotone, so they are applied to the reduced values in the order ``rescale_data_arr`` / ``get_ita``
        apply them to the arrays -- the results are the same floating-point numbers."""
        from .. import _lib
        from ..sim.kwave_if import _device
        out = SolutionAnalysis()
        units = options.distance_units
        dt = 1 / (self.pulse.frequency * 20)
        input_signal_V = self.pulse.calc_pulse(self.pulse.calc_time(dt)) * self.voltage
        res = self.simulation_result
        p_da, i_da = res["p_min"], res["intensity"]
        if list(p_da.dims)[0] != "focal_point_index" or p_da.ndim != 4:
            raise ValueError("the device analysis expects fields of dims (focal_point_index, x, y, z)")
        pnp_raw = np.asarray(p_da.data)
        ipa_raw = np.asarray(i_da.data)
        if pnp_raw.dtype != np.float32:
            raise ValueError(f"the device analysis expects the solver's float32 p_min, got {pnp_raw.dtype}")
        if ipa_raw.dtype != np.float64:
            ipa_raw = ipa_raw.astype(np.float64)
        pnp_scale = np.float32(getunitconversion(p_da.attrs["units"], "MPa"))
        ipa_scale = getunitconversion(i_da.attrs["units"], "W/cm^2")
        ita_scale = getunitconversion(i_da.attrs["units"], "mW/cm^2")
        dc_pt, dc_seq = self.get_pulsetrain_dutycycle(), self.get_sequence_dutycycle()
        if min(ipa_scale, ita_scale, dc_pt, dc_seq, float(pnp_scale)) < 0:
            raise ValueError("negative unit factor")

        def to_ipa(v):      # rescale_data_arr(intensity, "W/cm^2")
            return v * ipa_scale

        def to_ita(v):      # get_ita: rescale to mW/cm^2, then the two duty cycles, left to right
            return v * ita_scale * dc_pt * dc_seq

        if options.sidelobe_radius is np.nan:
            options.sidelobe_radius = options.mainlobe_radius
        dims = list(p_da.dims)[1:]
        axes, to_mm_axes = [], []
        for d in dims:
            c = p_da.coords[d]
            cu = c.attrs.get("units", None)
            a = np.asarray(c.data, dtype=np.float64)
            axes.append(getunitconversion(cu, units) * a if cu is not None else a)     # rescale_coords
            to_mm_axes.append(getunitconversion(units, "mm"))
        z_ok = axes[2] > options.sidelobe_zmin
        ar = options.mainlobe_aspect_ratio

        standoff_Z = options.standoff_density * 1500
        c_tic = 40e-3
        d_eq_cm = np.sqrt(4 * self.transducer.get_area("cm") / np.pi)
        ele_sizes_cm2 = np.array([el.get_area("cm") for el in self.transducer.elements])
        out.duty_cycle_pulse_train_pct = dc_pt * 100
        out.duty_cycle_sequence_pct = dc_seq * 100
        seq = self.sequence
        if seq.pulse_train_interval == 0:
            out.sequence_duration_s = float(seq.pulse_interval * seq.pulse_count * seq.pulse_train_count)
        else:
            out.sequence_duration_s = float(seq.pulse_train_interval * seq.pulse_train_count)

        nf = self.num_foci()
        power_W = np.zeros(nf)
        TIC = np.zeros(nf)
        to_mm = getunitconversion(units, "mm")
        global_all = np.nan
        stack = getattr(self, "_stack", None)
        with _lib.BeamAnalysis(axes, nf, z_ok=z_ok, device=_device() if stack is None else stack.device) as ana:
            for i in range(nf):
                if stack is not None:
                    (_, d_pnp, d_ipa), strides = stack.pointers(i)       # device -> device, no host staging
                    ana.set_focus(i, d_pnp, d_ipa, strides=strides)
                else:
                    ana.set_focus(i, pnp_raw[i], ipa_raw[i])
            for i in range(nf):
                focus = self.foci[i].get_position(units=units)
                focus_mm = self.foci[i].get_position(units="mm")
                out.target_position_lat_mm += [focus_mm[0]]
                out.target_position_ele_mm += [focus_mm[1]]
                out.target_position_ax_mm += [focus_mm[2]]
                apod = self.apodizations[i]
                origin = self.transducer.get_effective_origin(apodizations=apod, units=units)
                p0_Pa = np.max(self.transducer.calc_output(input_signal_V.copy(), dt, delays=self.delays[i, :], apod=apod), axis=1)

                frame = FocusFrame(focus, origin)
                offs, pts = [], []
                for k, scale in enumerate(ar):
                    o = np.linspace(-scale * options.beamwidth_radius, scale * options.beamwidth_radius, pnp_raw.shape[1 + k] * 2)
                    offs.append(o)
                    pts.append(frame.line(k, o))
                m, lines = ana.run_focus(i, frame.inverse, ar, options.mainlobe_radius, options.sidelobe_radius,
                                         pnp_scale, line_pts=pts)
                pk = m["main_pnp"]
                with np.errstate(invalid="ignore", divide="ignore"):
                    tot = np.float32(m["cen_w"])          # the reference's total is a float32 sum
                    cen = [float(np.float64(m[key]) / tot * s) for key, s in zip(("cen_wx", "cen_wy", "cen_wz"), to_mm_axes)]
                out.focal_centroid_lat_mm += [cen[0]]
                out.focal_centroid_ele_mm += [cen[1]]
                out.focal_centroid_ax_mm += [cen[2]]
                out.mainlobe_pnp_MPa += [pk]
                for k, named in enumerate(("lat", "ele", "ax")):
                    for db in (3, 6):
                        neg, pos = _bounds_from_line(offs[k], lines[k], pk * 10 ** (-db / 20))
                        name = f"beamwidth_{named}_{db}dB_mm"
                        setattr(out, name, [*getattr(out, name), to_mm * (pos - neg)])
                out.mainlobe_isppa_Wcm2 += [float(to_ipa(m["main_ipa"]))]
                out.mainlobe_ispta_mWcm2 += [float(to_ita(m["main_ipa_all"]))]
                side_pnp, side_ipa = m["side_pnp"], float(to_ipa(m["side_ipa"]))
                out.sidelobe_pnp_MPa += [side_pnp]
                out.sidelobe_isppa_Wcm2 += [side_ipa]
                out.sidelobe_to_mainlobe_pressure_ratio += [_ratio(side_pnp, out.mainlobe_pnp_MPa[-1])]
                out.sidelobe_to_mainlobe_intensity_ratio += [_ratio(side_ipa, out.mainlobe_isppa_Wcm2[-1])]
                out.global_pnp_MPa += [m["global_pnp"]]
                out.global_isppa_Wcm2 += [float(to_ipa(m["global_ipa"]))]
                global_all = m["global_ipa_all"]

                i0ta_Wcm2 = (p0_Pa ** 2 / (2 * standoff_Z)) * 1e-4 * out.duty_cycle_sequence_pct / 100
                power_W[i] = np.mean(np.sum(i0ta_Wcm2 * ele_sizes_cm2 * self.apodizations[i, :]))
                TIC[i] = power_W[i] / (d_eq_cm * c_tic)
                out.p0_MPa += [1e-6 * float(np.max(p0_Pa))]

        # nanmax(ita * z_ok): planes with z_ok False contribute ita * 0 = 0 (NaN where ita is not finite)
        cands = [to_ita(global_all)]
        if not bool(np.all(z_ok)):
            cands.append(0.0)
        cands = [c for c in cands if not np.isnan(c)]
        out.global_ispta_mWcm2 = float(max(cands)) if cands else float("nan")
        out.MI = float(np.max(out.mainlobe_pnp_MPa) / np.sqrt(self.pulse.frequency * 1e-6))
        out.TIC = float(np.mean(TIC))
        out.voltage_V = self.voltage
        out.power_W = float(np.mean(power_W))
        out.param_constraints = {} if param_constraints is None else param_constraints
        return out

    def compute_scaling_factors(self, focal_pattern: FocalPattern, analysis: SolutionAnalysis) -> Tuple[np.ndarray, float, float]:
        """(apodization factors per focus, old voltage, new voltage) that bring every focus'
        main-lobe PNP to the pattern's target pressure."""
        target_MPa = focal_pattern.target_pressure * getunitconversion(focal_pattern.units, "MPa")
        factors = np.array([target_MPa / analysis.mainlobe_pnp_MPa[i] for i in range(self.num_foci())], dtype=np.float64)
        top = np.max(factors)
        v0 = self.voltage
        return factors / top, v0, v0 * top

    def scale(self, focal_pattern: FocalPattern, analysis_options: SolutionAnalysisOptions = SolutionAnalysisOptions()) -> None:
        """Rescale apodizations, voltage and the stored fields in place to the target pressure."""
        analysis = self.analyze(options=analysis_options)
        apod_factors, v0, v1 = self.compute_scaling_factors(focal_pattern, analysis)
        stack = getattr(self, "_stack", None)
        for i in range(self.num_foci()):
            s = v1 / v0 * apod_factors[i]
            if stack is not None:
                stack.scale(i, s, s ** 2)    # same IEEE operations on the device-resident fields (csrc/stack.cu)
            else:
                self.simulation_result["p_min"][i].data *= s
                self.simulation_result["p_max"][i].data *= s
                self.simulation_result["intensity"][i].data *= s ** 2
            self.apodizations[i] = self.apodizations[i] * apod_factors[i]
        self.voltage = v1

    def get_pulsetrain_dutycycle(self) -> float:
        return min(1., self.pulse.duration / self.sequence.pulse_interval)

    def get_sequence_dutycycle(self) -> float:
        seq = self.sequence
        between = 1 if seq.pulse_train_interval == 0 else (seq.pulse_count * seq.pulse_interval) / seq.pulse_train_interval
        return self.get_pulsetrain_dutycycle() * between

    def get_ita(self, units: str = "mW/cm^2"):
        """Time-averaged intensity per focus: intensity x pulse-train duty cycle x sequence duty cycle
        (the value the reference's count-weighted expression reduces to, solution.py:365-388)."""
        ita = rescale_data_arr(self.simulation_result["intensity"], units).copy(deep=True)
        ita.data = np.asarray(ita.data, dtype=np.float64) * self.get_pulsetrain_dutycycle() * self.get_sequence_dutycycle()
        return ita

    # ------------------------------------------------------------------------------ (de)serialisation
    def to_dict(self, include_simulation_data: bool = False) -> dict:
        d = asdict(self)
        if not include_simulation_data:
            d.pop("simulation_result")
        return d

    def _plain_dict(self) -> dict:
        """The dictionary the reference serialises (``dataclasses.asdict(self)``, plan/solution.py:416): nested
        dataclasses field by field, ``None`` members included, so that files written here and by the reference have
        the same layout (tests/golden/ref_solution.json is a reference-written file)."""
        d = asdict(replace(self, simulation_result=None))
        d.pop("simulation_result")
        return d

    def to_json(self, include_simulation_data: bool, compact: bool) -> str:
        d = self._plain_dict()
        if include_simulation_data:
            with tempfile.NamedTemporaryFile(suffix=".nc", delete=False) as tmp:
                tmp_path = Path(tmp.name)
            try:
                self.simulation_result.to_netcdf(tmp_path, engine="scipy")
                raw = tmp_path.read_bytes()
            finally:
                tmp_path.unlink(missing_ok=True)
            d["simulation_result"] = base64.b64encode(raw).decode("utf-8")
        if compact:
            return json.dumps(d, separators=(",", ":"), cls=_Encoder)
        return json.dumps(d, indent=4, cls=_Encoder)

    @staticmethod
    def from_dict(solution_dict: dict) -> "Solution":
        d = dict(solution_dict)
        if isinstance(d.get("date_created"), str):
            d["date_created"] = datetime.fromisoformat(d["date_created"])
        if d.get("delays") is not None:
            d["delays"] = np.array(d["delays"])
        if d.get("apodizations") is not None:
            d["apodizations"] = np.array(d["apodizations"], ndmin=2)
        if d.get("transducer") is not None and not isinstance(d["transducer"], Transducer):
            d["transducer"] = Transducer.from_dict(d["transducer"])
        if not isinstance(d.get("pulse", Pulse()), Pulse):
            d["pulse"] = Pulse.from_dict(d["pulse"])
        if not isinstance(d.get("sequence", Sequence()), Sequence):
            d["sequence"] = Sequence.from_dict(d["sequence"])
        d["foci"] = [p if isinstance(p, Point) else Point.from_dict(p) for p in d.get("foci", [])]
        if d.get("target") is not None and not isinstance(d["target"], Point):
            d["target"] = Point.from_dict(d["target"])
        if isinstance(d.get("simulation_result"), str):
            raw = base64.b64decode(d["simulation_result"].encode("utf-8"))
            d["simulation_result"] = xa.open_dataset(raw, engine="scipy")
        return Solution(**d)

    @staticmethod
    def from_json(json_string: str, simulation_result=None) -> "Solution":
        d = json.loads(json_string)
        if simulation_result is not None:
            if "simulation_result" in d:
                raise ValueError("A simulation result was provided while the json string already contains "
                                 "`simulation_result`. Unclear which to use!")
            d["simulation_result"] = simulation_result
        return Solution.from_dict(d)

    def to_files(self, json_filepath: Path, nc_filepath: Path | None = None) -> None:
        """JSON + netCDF pair (reference ``to_files``, solution.py:499-516).  The reference writes netCDF-4 through
        ``engine='h5netcdf'``; where that backend is not installed (h5py / h5netcdf are optional and absent from the
        offline image) the fields are written as NetCDF-3 64-bit-offset (``engine='scipy'``) instead -- announced at
        WARNING level; ``from_files`` recognises either container by its magic bytes, and xarray itself opens both."""
        json_filepath = Path(json_filepath)
        nc_filepath = _nc_path_for(json_filepath) if nc_filepath is None else Path(nc_filepath)
        json_filepath.parent.mkdir(parents=True, exist_ok=True)
        nc_filepath.parent.mkdir(parents=True, exist_ok=True)
        json_filepath.write_text(self.to_json(include_simulation_data=False, compact=False))
        if _netcdf4_available():
            self.simulation_result.to_netcdf(nc_filepath, engine="h5netcdf")
        else:
            logging.getLogger(__name__).warning(
                "Solution.to_files: no netCDF-4 backend (h5netcdf / h5py) is installed; writing %s as NetCDF-3 "
                "(64-bit offset, engine='scipy')", nc_filepath)
            self.simulation_result.to_netcdf(nc_filepath, engine="scipy")

    @staticmethod
    def from_files(json_filepath: Path, nc_filepath: Path | None = None) -> "Solution":
        json_filepath = Path(json_filepath)
        nc_filepath = _nc_path_for(json_filepath) if nc_filepath is None else Path(nc_filepath)
        with open(nc_filepath, "rb") as f:
            magic = f.read(4)
        engine = "scipy" if magic[:3] == b"CDF" else "h5netcdf"        # NetCDF-3 classic / 64-bit offset vs HDF5
        ds = xa.open_dataset(nc_filepath, engine=engine).load()
        ds.close()
        return Solution.from_json(json_filepath.read_text(), simulation_result=ds)


def _netcdf4_available() -> bool:
    """Is there a backend that writes netCDF-4 (HDF5) files?  Needs real xarray with h5netcdf + h5py."""
    if not getattr(xa, "HAVE_XARRAY", False):
        return False
    try:
        import h5netcdf  # noqa: F401
        import h5py  # noqa: F401
    except ImportError:
        return False
    return True

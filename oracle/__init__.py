"""CPU oracle for the OpenLIFU treatment-planning hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and there only as the checker / CPU baseline.  The product
package (``openlifu-python_b200/``) never imports ``oracle``.

What it restates (file:line relative to /root/reference):

* the beamforming inputs -- ``bf/delay_methods/direct.py:28-38``,
  ``bf/apod_methods/{uniform,maxangle,piecewiselinear}.py``,
  ``bf/focal_patterns/{single,wheel}.py``, ``geo.py:56-74``,
  ``xdc/element.py:144-260``, ``xdc/transducer.py:95-112,372-406``
  (module ``oracle.beamform``).  PINNED: checked against golden vectors produced
  by importing the real reference in the build container
  (``tests/golden/make_reference_goldens.py`` -> ``tests/golden/ref_beamform.npz``).
* the adapter ``sim/kwave_if.py`` (entire file) and the grid definition
  ``sim/sim_setup.py:54-116,152-155`` (modules ``oracle.scene``, ``oracle.kgrid``).
* the third-party solver the adapter delegates to: ``k-wave-python==0.4.0``
  (``pyproject.toml:46``) and the ``kspaceFirstOrder-OMP`` binary it downloads
  (``src/openlifu/util/assets.py:134-154``).  That source is NOT under
  /root/reference and cannot be installed here (no network), so modules
  ``oracle.kgrid`` / ``oracle.bli`` / ``oracle.solver`` restate its published
  algorithm (k-Wave manual; Treeby & Cox JBO 2010; Treeby et al. JASA 2012;
  Wise et al. JASA 2019) anchored on the reference call sites
  ``sim/kwave_if.py:13-27,29-47,49-63,65-78,117-129``.

PARITY UNPINNED for the solver boundary: the reference's only test of this path
(``tests/test_sim.py:18-60``) asserts types and keys, not numbers, and no k-Wave
binary or golden field exists in the container.  The solver restatement is
therefore checked against analytic known answers instead (plane-wave phase
velocity, free-space Green's function, PML decay, power-law absorption decay --
``tests/test_oracle_physics.py``).  Every assumption about k-Wave internals is a
named switch in ``oracle.solver.Assumptions`` (ledger A1-A11 of SURVEY.md 8c).
"""

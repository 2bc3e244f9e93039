"""
oracle/ -- CPU restatement of the reference hot path.  TEST INFRASTRUCTURE ONLY.

Nothing under ``glass_b200/`` may import, call, link or execute anything in this
package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` use it, and there only as the checker
(or as the timed CPU arm), never as the product path.

Parity status (see DESIGN.md "Oracle"):

* GLASS-side numpy code (``glass/fields.py``, ``points.py``, ``lensing.py``,
  ``galaxies.py``, ``shapes.py``, ``harmonics.py``, ``grf/_transformations.py``):
  restated in :mod:`oracle.glass_ref`; **pinned** by golden vectors produced by
  executing the reference's own source in the build container
  (``tests/golden/make_golden.py``) and by the reference's dependency-free
  known-answer tests, ported in ``tests/test_oracle_*.py``.
* Third-party leaf maths reached through ``glass/healpix.py`` (healpy /
  libsharp2 ``alm2map``, ``alm2map_spin``, ``map2alm``; ``healpix.randang``,
  ``ang2pix``): the libraries are absent from ``/root/reference`` and from this
  image and are not version-pinned by the reference (``pyproject.toml:80-86``
  gives lower bounds only: healpy>=1.15.0, healpix>=2022.11.1).  The published
  algorithms are restated in :mod:`oracle.healpix_ref` (numpy) and
  ``oracle/sht_ref.c`` (C) and validated against the mathematical definition
  (direct summation of ``scipy.special.sph_harm_y`` / Wigner-d spin harmonics at
  pixel centres).  The reference's own tests for these functions are purely
  differential against the live library (``tests/core/test_healpix.py``), hold
  no stored vectors, so for these functions: **parity unpinned**.
* ``oracle/transformcl_ref.py``: restatement of the absent third-party ``transformcl`` pair that
  the reference's spectra solver calls; it is the module under which ``make_golden.py --solver``
  executes the reference's own solver source (so the solver LOGIC is pinned by the reference),
  and the independent checker of ``glass_b200.transformcl``.  transformcl itself: unpinned.
* ``oracle/sht_fast.cpp`` is not a checker but the TIMED CPU arm of ``bench.py`` (a SIMD /
  OpenMP synthesis with the structure of libsharp, so that the CPU baseline is a fair one); it
  is itself pinned against ``sht_ref.c`` in ``tests/test_cpu_oracle.py``.
"""

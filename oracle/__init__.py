"""
CPU oracle for the Nemo per-tile filter-and-detect hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE.  It is a numpy/scipy restatement of the reference's own
arithmetic (``nemo/filters.py``, ``nemo/photometry.py``, parts of ``nemo/maps.py`` and
``nemo/signals.py``), function for function, in fp64.  Only ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it, and only as the checker
or the timed CPU baseline -- never as the product path.  ``nemo_b200`` never imports it.

PARITY STATUS: **pinned against runs of the reference's own code** for the Fourier matched-filter path, the noise map
(grid, 'smart' and 'percentile' branches), the S/N / mask block, ``getObjectPositions``, ``findObjects``, ``measureFluxes`` and
the object names; **unpinned** for the restated
third-party pieces and for what needs them.  The reference cannot be imported as a package in this image (astLib, astropy,
pixell, pyccl, mahotas, mpi4py absent; this fork even assigns ``enmap = None``, nemo/filters.py:22-27) and ships no golden
vectors or numeric known-answer tests for this path (tests/ is Robot-Framework integration tests on downloaded data).  But its
source files do run when the absent packages are replaced by stand-ins: ``tests/golden/make_reference_golden.py`` imports
``nemo/filters.py`` and ``nemo/photometry.py`` unmodified from ``/root/reference``, gives them the pixell / astLib functions
restated in ``pixell_shim.py`` and ``wcs_stub.py`` and the template maps as inputs, and stores what ``filters.filterMaps``
(``MapFilter.__init__``, ``makeNoiseMap``, ``MatchedFilter.buildAndApply`` / ``applyFilter``) and
``photometry.getObjectPositions / findObjects / measureFluxes`` (with ``catalogs.catalogListToTab / makeName``; a plain column
table stands in for ``astropy.table``) compute for the golden tiles in ``tests/golden/ref_*.npz``.  The oracle reproduces those
fixtures BIT FOR BIT (``tests/test_oracle.py::test_oracle_matches_reference_run``: filtered map, S/N map, area mask,
normalisation, relative weights, labels, positions, and every column of the catalogs), and the CUDA path is tested against the same fixtures
(``tests/test_golden_gpu.py``).  Still restated from published behaviour only: pixell ``enmap.fft/ifft/apod/apply_window`` and the
astLib CAR WCS and J2000 -> Galactic rotation (pinned by the known-answer tests in ``tests/test_oracle.py``), the
signal-template painters and the real-space filter's kernel construction (it deletes the in-memory maps in this fork,
nemo/filters.py:919-923, so it cannot run on arrays; the filter's ``applyFilter``, ``catalogs.makeOptimalCatalog`` and
``maps.getPixelAreaArcmin2Map`` do run and are pinned by ``tests/golden/ref_misc.npz``).  The scipy / numpy calls the reference makes (``ndimage.label``, ``center_of_mass``,
``gaussian_filter``, ``rank_filter``, ``convolve``, ``RectBivariateSpline``, ``np.linalg.inv``, ``np.percentile``) are called
here for real, so those are the reference's arithmetic exactly.
"""

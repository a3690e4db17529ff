"""CPU oracle for the HALO hyperbolic head + acquisition hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``halo_b200/`` may import this package.
Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may call it, and there only as the checker or the timed
CPU baseline -- never as the thing shipped.

What it is: a plain PyTorch-CPU / numpy restatement of the reference algorithm,
function by function, each citing the reference file:line it follows
(paths relative to the reference checkout):

* ``geoopt_math``  -- third-party ``geoopt.manifolds.stereographic.math`` (NOT vendored by
  the reference, version unpinned in ``requirements.txt:15``): expmap0 / project / dist0.
* ``head``         -- ``core/utils/hyperbolic.py:28-39,74-83,100-188`` (float64, as the reference).
* ``score``        -- ``core/active/floating_region.py:12-23,26-217``.
* ``select``       -- ``core/active/build.py:27-64`` (sequential greedy arg-max selection).
* ``acquire``      -- ``core/active/build.py:71-160`` minus model forward / file I/O.
* ``delta``        -- ``core/active/build.py:45-48,58-62``: the window labelling of a round, as exchanged between
  image shards by ``halo_round_delta_pack`` / ``halo_round_delta_apply``.

Parity pinning: the reference ships no tests, golden vectors or fixtures for this path
("parity unpinned" by the reference's own tests).  The oracle is
therefore pinned by running the reference's OWN modules (imported unchanged from the
reference checkout behind stubs for the absent third-party packages, see
``oracle/ref_import.py``) on seeded inputs: ``tests/golden/make_golden.py`` freezes those
outputs as fixtures under ``tests/golden/`` and ``tests/test_oracle_golden.py``
re-checks the restatement against the live reference whenever the checkout is present.
The only part that cannot be pinned against real code is the geoopt boundary (geoopt is
not installed and cannot be fetched); its restatement follows geoopt's published
``stereographic/math.py`` (k<0 branch) and its guards (1e-15 / 1e-7 / clamp 15) influence
results by <= 1e-14 relative.
"""

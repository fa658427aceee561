"""CPU oracle for the DistDiff hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is product code.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` leg may import it, and only as the checker / CPU baseline; the
product package ``distdiff_b200`` never imports it and fails loudly when its
CUDA extension is missing.

The oracle is a literal fp32 (fp64 where the reference's third-party code is
fp64) restatement of the reference's algorithm, every function citing the
``/root/reference`` file:line it follows:

* ``oracle.ddim``        generate_data.py:109-121, 1043-1044, 1170-1180, 1002-1009
                         + the restated diffusers ``DDIMScheduler`` contract
* ``oracle.energy``      generate_data.py:707-717 / 747-759 (+ analytic gradient)
* ``oracle.guidance``    generate_data.py:124-137, 687-732, 735-767
* ``oracle.prototypes``  dataloader.py:664-731, generate_data.py:1113-1127
                         (agglomerative = the reference's own sklearn/scipy call;
                         k-means = north-star extension, parity unpinned by the
                         reference)

Pinning: the reference has no tests, golden vectors or fixtures for this path
(SURVEY.md section 4).  The oracle is pinned instead against OUTPUTS OF THE
REFERENCE'S OWN FUNCTION BODIES: ``tests/golden/make_golden.py`` extracts
``denoise_one_step``, ``tensor_clamp``, ``linfball_proj``,
``transform_guidance``, ``direct_guidance`` (generate_data.py) and
``extract_prototype`` (dataloader.py) from the reference sources with ``ast``,
executes them unmodified on CPU with stubbed third-party objects, and commits
the resulting vectors under ``tests/golden/``.  What stays restated (diffusers
is not installable here, and is unpinned upstream) is the DDIM scheduler
arithmetic; it is pinned by closed-form identities and the alpha-bar values
verified in SURVEY.md section 8c.  The k-means path has no reference
counterpart at all: "parity unpinned" -- its oracle is the spec.
"""

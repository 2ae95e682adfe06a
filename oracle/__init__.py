"""CPU oracle for the NeRF volume-rendering hot path -- TEST INFRASTRUCTURE ONLY.

Nothing under ``oracle/`` is part of the product.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import it, and only as the checker / the timed CPU baseline.  The product
path (``smpl_nerf_b200``) never imports this package and fails loudly if its CUDA
library is missing.

Contents
--------
nerf_oracle.py        torch-CPU (fp32 or fp64) restatement of the reference hot path
                      (utils.py:114-264, models/render_ray_net.py:42-61,
                      models/warp_field_net.py:17-21, models/*_pipeline.py forward).
                      Pinned bit-for-bit against the imported reference in the build
                      container (tests/test_oracle_vs_reference.py) and against the
                      committed golden fixtures in tests/golden/ everywhere else.
ref_import.py         imports the *actual* reference modules from /root/reference with
                      sys.modules stubs; only usable where that
                      path exists (the build container, never the GPU box).
searchsorted_oracle.c plain-C restatement of torchsearchsorted's bisection
                      (torchsearchsorted/src/cpu/searchsorted_cpu_wrapper.cpp:4-122).
scene.py              synthetic SMPL-NeRF scene (cameras, coarse sampling, poses) that
                      mirrors create_dataset.py defaults.
Makefile              builds searchsorted_oracle.c -> oracle/_build/ and, when
                      /root/reference is present, the reference's own C++ CPU
                      searchsorted -> oracle/_ref/ (never copied into the repo).
"""

"""CPU oracle for the SuRS reconstruction hot path -- TEST INFRASTRUCTURE ONLY.

Nothing in the product package (``surs_b200`` / ``super-resolution-...-image_b200``)
imports this package.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may use it, and there
only as the checker / CPU baseline, never as the thing shipped.

Contents
--------
``surs_oracle``   numpy float64 restatement of query / create_grid / eval_grid /
                  eval_grid_octree (pinned against the reference's own modules run
                  in the build container -> ``tests/golden/*.npz``).
``mc_oracle``     ctypes loader for ``mc_oracle.c`` -- a scalar C marching cubes
                  (PARITY UNPINNED vs scikit-image 0.17.2, see its header).
``ref_import``    imports the unmodified reference from /root/reference (build
                  container only) to generate the golden fixtures.
"""

"""CPU oracle for the RollingDepth snippet-inference hot path.

TEST INFRASTRUCTURE ONLY.  Nothing under ``rollingdepth_b200/`` may import this
package; only ``tests/``, ``__graft_entry__.smoke()`` and the ``cpu_baseline`` /
``--impl reference`` legs of ``bench.py`` do, and there only as the checker or the
timed CPU baseline -- never as the product path.

Every function restates one piece of the reference (yizuo417/RollingDepth,
mounted read-only at /root/reference while this repo was built) and cites the
file:line it follows.  The reference has NO tests or golden vectors of its own
for the cross-frame attention, the pipeline or the aligner;
parity is pinned instead by fixtures generated from the reference's own code
imported in the build container (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``), and by the reference's DDIM known-answer values
(diffusers/tests/schedulers/test_scheduler_ddim.py:114-176).
"""

"""TEST INFRASTRUCTURE ONLY.

CPU oracle for the deep-fluids generator/AE forward-backward hot path.  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s `cpu_baseline` / `--impl reference` legs may import this
package -- and there only as the checker / the timed CPU baseline, never as the shipped path.
See oracle/ref_ops.py for the pinning status (stencils pinned against the reference's own ops.py;
conv/FC/Adam "parity unpinned": TensorFlow 1.15 is absent and the reference has no tests).
"""

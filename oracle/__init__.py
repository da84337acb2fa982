"""oracle/ -- CPU restatement of pyratslam's hot path.  TEST INFRASTRUCTURE ONLY.

This package restates, in Python 3 + numpy/scipy float64, what the reference
(bjkomer/pyratslam, Python 2 + OpenCL) computes on its hot path:

    oracle.posecells       <- ratslam/posecell_network.py + ratslam/convolution.py
    oracle.view_templates  <- ratslam/view_templates.py
    oracle.experience_map  <- ratslam/experience_map.py
    oracle.drivers         <- loop order of ratslam/simulate.py and ratslam/ros_simulate.py

It is the *checker*, never the product: only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl
reference`` legs may import it.  Nothing under ``pyratslam_b200/`` imports it,
and the product path raises if the CUDA library is missing instead of falling
back to this code.

Parity pin
----------
The reference ships no tests, golden vectors or published numbers, and its
device path (pyopencl + mako under Python 2) cannot be executed in the build
container.  The pin used instead: ``tests/golden/make_golden.py`` executes the
reference's *own* host code (``posecell_network.py``, ``convolution.py``,
``view_templates.py``, ``experience_map.py`` read from ``/root/reference`` at
generation time, never copied) under a Python-2 semantics shim, with the three
OpenCL kernels on the path emulated by their literal flat-index arithmetic in
numpy.  The fixtures it wrote live in ``tests/golden/*.npz``; the ``-m "not
gpu"`` suite checks this oracle against every one of them.  What remains
unpinned is only the OpenCL runtime's own ``double`` rounding (FMA contraction
is implementation-defined there), which the reference author's sandbox check
(``sandbox/opencl_test2.py:306-334``) equates with ``scipy.ndimage``'s.
"""

"""Run the reference's own Python-2 host code under Python 3 -- fixture generator support.

Only ``make_golden.py`` (run once, in the build container, where
``/root/reference`` is mounted) imports this.  Nothing here is copied from the
reference: its sources are *read from /root/reference at run time*, adapted to
Python 3 by the mechanical, documented rewrites below, and executed.

Rewrites (each one restores Python-2 / numpy-1.x behaviour, none changes an
algorithm):

R1  ``raise X, "msg"``                       -> ``raise X("msg")``       (py2 syntax)
R2  ``a / b``                                -> ``_py2div(a, b)``         (floor division when both
                                                operands are integers or integer arrays, as in py2)
R3  ``from numpy import *``                  -> star-import of a shim that exports what numpy 1.x
                                                exported: every public numpy name **except**
                                                ``abs/max/min/round`` (builtins survive) **plus**
                                                ``math`` (numpy.lib re-exported it)
R4  ``xrange``                               -> ``range``;  ``math.floor/ceil`` return floats (py2)
R5  ``self.radius = radius`` (convolution.py:632) -> ``int(radius)``: numpy 1.x accepted the float
                                                ``ceil()`` result as a slice bound, numpy 2.x refuses.

Stand-ins for packages absent from the container:

* ``mako.template.Template``  -- ``${name}`` substitution only (all the reference uses).
* ``pyopencl``                -- contexts/queues/buffers as numpy arrays; ``Program.build()`` returns
  an object whose ``conv``, ``conv_xy_origin_filters`` and ``conv_z`` methods evaluate the OpenCL-C
  kernels of ``convolution.py:228-246,320-340,344-359`` by their **literal flat-index arithmetic**
  over the padded buffers (one numpy gather per filter tap, taps accumulated in the kernels' loop
  order, multiply and add rounded separately).
"""
from __future__ import annotations

import ast
import math
import re
import sys
import types

import numpy as np

REF_DIR = "/root/reference/ratslam"


# --------------------------------------------------------------------------- R2
def _is_intlike(v):
    if isinstance(v, (bool, np.bool_)):
        return False
    if isinstance(v, (int, np.integer)):
        return True
    return isinstance(v, np.ndarray) and np.issubdtype(v.dtype, np.integer)


def _py2div(a, b):
    if _is_intlike(a) and _is_intlike(b):
        return a // b
    return a / b


class _DivRewriter(ast.NodeTransformer):
    def visit_BinOp(self, node):
        self.generic_visit(node)
        if isinstance(node.op, ast.Div):
            return ast.copy_location(
                ast.Call(func=ast.Name(id="_py2div", ctx=ast.Load()), args=[node.left, node.right], keywords=[]),
                node)
        return node


# --------------------------------------------------------------------------- R1
def _fix_raise(src):
    # join backslash continuations so a multi-line raise is one logical line
    joined = re.sub(r"\\\n\s*", " ", src)
    out = []
    for line in joined.split("\n"):
        m = re.match(r"^(\s*)raise\s+(\w+)\s*,\s*(.+)$", line)
        if m:
            line = "%sraise %s(%s)" % (m.group(1), m.group(2), m.group(3))
        out.append(line)
    return "\n".join(out)


# --------------------------------------------------------------------------- R3
def _py2_math():
    """``math`` as Python 2 had it: ``floor``/``ceil`` return floats (so ``x - center`` and the
    divisions that follow stay floating point, e.g. posecell_network.py:101,109-110)."""
    m = types.ModuleType("math")
    m.__dict__.update({k: v for k, v in math.__dict__.items() if not k.startswith("__")})
    m.floor = lambda x: float(math.floor(x))
    m.ceil = lambda x: float(math.ceil(x))
    return m


def _numpy1_star_module():
    m = types.ModuleType("_np_py2_star")
    names = [n for n in dir(np) if not n.startswith("_") and n not in ("abs", "max", "min", "round")]
    for n in names:
        try:
            setattr(m, n, getattr(np, n))
        except Exception:
            pass
    m.math = _py2_math()
    m.__all__ = [n for n in names if hasattr(m, n)] + ["math"]
    return m


# --------------------------------------------------------------------------- mako stand-in
class _Rendered(str):
    conf = None


class _Template:
    def __init__(self, text, output_encoding=None):
        self.text = text

    def render(self, **conf):
        out = re.sub(r"\$\{(\w+)\}", lambda m: str(conf[m.group(1)]), self.text)
        r = _Rendered(out)
        r.conf = dict(conf)
        return r


# --------------------------------------------------------------------------- pyopencl stand-in
class _Buf:
    def __init__(self, ctx, flags, size=None, hostbuf=None):
        if hostbuf is not None:
            self.data = np.array(hostbuf).ravel().copy()      # COPY_HOST_PTR
        else:
            self.data = np.zeros(size // 8, dtype=np.float64)  # only float64 images on this path


class _Event:
    def wait(self):
        return None


class _Program:
    def __init__(self, ctx, text):
        self.conf = text.conf

    def build(self):
        return self

    # convolution.py:228-246 ------------------------------------------------
    def conv(self, queue, gshape, lshape, im, fil, out):
        c = self.conf
        fs, off, len_y, len_z = c["filsize"], c["offset"], c["len_y"], c["len_z"]
        I, J, K = np.meshgrid(*[np.arange(n) for n in gshape], indexing="ij")
        acc = np.zeros(gshape, dtype=im.data.dtype)
        for x in range(fs):
            for y in range(fs):
                for z in range(fs):
                    idx = ((off + K + z - off) + (off + J + y - off) * len_z
                           + (off + I + x - off) * len_z * len_y)
                    acc = acc + im.data[idx] * fil.data[z + y * fs + x * fs * fs]
        out.data[off + K + (J + off) * len_z + (I + off) * len_z * len_y] = acc

    # convolution.py:320-340 ------------------------------------------------
    def conv_xy_origin_filters(self, queue, gshape, lshape, im, fil, out, origin_x, origin_y,
                               radius, len_y, len_z):
        c = self.conf
        fs, off = c["filsize"], c["offset"]
        radius, len_y, len_z = int(radius), int(len_y), int(len_z)
        I, J, K = np.meshgrid(*[np.arange(n) for n in gshape], indexing="ij")
        ox = origin_x.data.astype(np.int64)[K]
        oy = origin_y.data.astype(np.int64)[K]
        acc = np.zeros(gshape, dtype=im.data.dtype)
        for x in range(fs):
            for y in range(fs):
                idx = ((off + K) + (off + oy + radius + J + y - off) * len_z
                       + (off + ox + radius + I + x - off) * len_z * len_y)
                fidx = K + y * (len_z - 2 * off) + x * (len_z - 2 * off) * fs
                acc = acc + im.data[idx] * fil.data[fidx]
        out.data[off + K + (J + off) * len_z + (I + off) * len_z * (len_y - 2 * radius)] = acc

    # convolution.py:344-359 ------------------------------------------------
    def conv_z(self, queue, gshape, lshape, im, fil, out):
        c = self.conf
        fs, off, len_y, len_z = c["filsize"], c["offset"], c["len_y"], c["len_z"]
        I, J, K = np.meshgrid(*[np.arange(n) for n in gshape], indexing="ij")
        acc = np.zeros(gshape, dtype=im.data.dtype)
        for z in range(fs):
            idx = (off + K + z - off) + (off + J) * len_z + (off + I) * len_z * len_y
            acc = acc + im.data[idx] * fil.data[z]
        out.data[off + K + (J + off) * len_z + (I + off) * len_z * len_y] = acc


def _fake_pyopencl():
    cl = types.ModuleType("pyopencl")
    cl.create_some_context = lambda: object()
    cl.CommandQueue = lambda ctx: object()
    cl.mem_flags = types.SimpleNamespace(READ_ONLY=1, WRITE_ONLY=2, COPY_HOST_PTR=4, READ_WRITE=8)
    cl.Buffer = _Buf
    cl.Program = _Program

    def enqueue_read_buffer(queue, buf, out):
        out.ravel()[...] = buf.data[: out.size].astype(out.dtype, copy=False)
        return _Event()

    cl.enqueue_read_buffer = enqueue_read_buffer
    return cl


# --------------------------------------------------------------------------- loader
def _load(name, extra_patch=None):
    src = open("%s/%s.py" % (REF_DIR, name)).read()
    src = _fix_raise(src)
    src = src.replace("from numpy import *", "from _np_py2_star import *")
    if extra_patch:
        for old, new in extra_patch:
            assert old in src, (name, old)
            src = src.replace(old, new)
    tree = ast.parse(src, filename="%s/%s.py" % (REF_DIR, name))
    tree = ast.fix_missing_locations(_DivRewriter().visit(tree))
    mod = types.ModuleType(name)
    mod.__dict__["_py2div"] = _py2div
    mod.__dict__["xrange"] = range
    sys.modules[name] = mod
    exec(compile(tree, "%s/%s.py" % (REF_DIR, name), "exec"), mod.__dict__)
    return mod


def load_reference():
    """Returns the reference modules (posecell_network, convolution, view_templates, experience_map)."""
    sys.modules["_np_py2_star"] = _numpy1_star_module()
    sys.modules["pyopencl"] = _fake_pyopencl()
    mako = types.ModuleType("mako")
    mako_t = types.ModuleType("mako.template")
    mako_t.Template = _Template
    mako.template = mako_t
    sys.modules["mako"] = mako
    sys.modules["mako.template"] = mako_t
    conv = _load("convolution", extra_patch=[("self.radius = radius ", "self.radius = int(radius) ")])
    pcn = _load("posecell_network")
    vt = _load("view_templates")
    em = _load("experience_map")
    return types.SimpleNamespace(convolution=conv, posecell_network=pcn, view_templates=vt, experience_map=em)

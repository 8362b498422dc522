"""A minimal NumPy-backed stand-in for the slice of TensorFlow 1.13 that the reference's hot-path
modules touch, so that the reference's OWN source files (``/root/reference/models/*.py``) can be
imported and executed in the build container, where TensorFlow cannot be installed.

Used ONLY by ``make_golden.py`` (golden-vector generation); never by the product, the oracle, the
tests at run time or the bench.  What is pinned by running the reference through it: all the
Python-level arithmetic and wiring in ``entropy_model.py``, ``conditional_entropy_model.py``,
``model_voxception.py`` and ``model_simple.py`` (operation order, signs, floors, transposes, layer
graph, bias/activation flags).  What is NOT pinned (it lives inside TF itself, not in the reference
tree): the conv kernels (stood in for by torch-CPU with the SAME/transposed rules restated in
``oracle/nets.py``) and ``tf.contrib.coder`` (stood in for by ``oracle/coder.py``).
"""
from __future__ import annotations

import sys
import types

import numpy as np

CAPTURE = {}          # last pmf handed to pmf_to_quantized_cdf etc.
WEIGHTS = {}          # "<layer>/kernel" | "<layer>/bias" | bottleneck variable names -> arrays


def _np(x):
    return np.asarray(x)


class _Dim:
    def __init__(self, v):
        self.value = v


class TensorShape:
    def __init__(self, dims):
        if isinstance(dims, TensorShape):
            dims = dims.dims
        self.dims = [None if d is None else int(d) for d in dims]
        self.ndims = len(self.dims)

    def __getitem__(self, i):
        return _Dim(self.dims[i])


class InputSpec:
    def __init__(self, ndim=None, axes=None, min_ndim=None):
        self.ndim, self.axes, self.min_ndim = ndim, axes, min_ndim


class _Initializer:
    def __init__(self, kind, *a):
        self.kind, self.a = kind, a


class Layer:
    def __init__(self, *a, **k):
        self.built = False
        self._dtype = "float32"
        self.input_spec = None

    @property
    def dtype(self):
        return self._dtype

    def _name_scope(self):
        return "layer"

    def add_variable(self, name, dtype=None, shape=None, initializer=None):
        v = np.asarray(WEIGHTS[name], np.float32)
        assert tuple(v.shape) == tuple(int(s) for s in shape), (name, v.shape, shape)
        return v

    def build(self, input_shape):
        self.built = True

    def __call__(self, inputs, *args, **kwargs):
        if not self.built:
            self.build(np.shape(inputs))
        return self.call(inputs, *args, **kwargs)


class Model(Layer):
    def __init__(self, name=None):
        Layer.__init__(self)

    def __call__(self, *args, **kwargs):
        return self.call(*args, **kwargs)

    def summary(self):
        return ""


class _ConvBase:
    transposed = False

    def __init__(self, filters, kernel_size, strides=(1, 1, 1), padding="valid", activation=None, use_bias=True, name=None):
        assert padding == "same"
        self.filters, self.k, self.s = filters, kernel_size, strides
        self.activation, self.use_bias, self.name = activation, use_bias, name

    def __call__(self, x):
        # direct-definition NumPy convolutions (tests/golden/ref_conv.py): independent of oracle/nets.py, so the goldens pin the
        # SAME-padding / Conv3DTranspose arithmetic and not only the layer graph
        import ref_conv
        kern = WEIGHTS[self.name + "/kernel"]
        has_bias = (self.name + "/bias") in WEIGHTS
        assert has_bias == bool(self.use_bias), "use_bias mismatch for %s" % self.name
        bias = WEIGHTS[self.name + "/bias"] if has_bias else None
        assert tuple(kern.shape[:3]) == tuple(self.k)
        x = np.ascontiguousarray(x, dtype=np.float32)
        s = self.s[0]
        if self.transposed:
            assert kern.shape[3] == self.filters
            y = ref_conv.conv3d_transpose_same(x, kern, bias, stride=s)
        else:
            assert kern.shape[4] == self.filters
            y = ref_conv.conv3d_same(x, kern, bias, stride=s)
        if self.activation is not None:
            y = self.activation(y)
        return y


class Conv3D(_ConvBase):
    pass


class Conv3DTranspose(_ConvBase):
    transposed = True


def _sigmoid(x):
    x = _np(x)
    out = np.empty_like(x)
    pos = x >= 0
    out[pos] = 1 / (1 + np.exp(-x[pos]))
    e = np.exp(x[~pos])
    out[~pos] = e / (1 + e)
    return out


class _NameScope:
    def __init__(self, *a):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _pmf_to_quantized_cdf(pmf, precision=16):
    from oracle import coder
    CAPTURE["pmf"] = np.array(pmf, np.float32)
    return coder.pmf_to_quantized_cdf(np.asarray(pmf, np.float32), precision)


def _range_encode(values, cdf, precision=16):
    from oracle import coder
    values = np.asarray(values)
    cdf = np.asarray(cdf)
    rows = cdf.reshape(-1, cdf.shape[-1])
    n = values.size
    # broadcast of cdf leading dims against data: here either [1,C,N+1] vs [M,C] or [M,C,N+1] vs [M,C]
    if cdf.shape[0] == 1:
        idx = np.tile(np.arange(rows.shape[0], dtype=np.int32), n // rows.shape[0])
    else:
        idx = np.arange(n, dtype=np.int32)
    return coder.range_encode(values.reshape(-1), rows, idx, precision)


def _range_decode(strings, shape, cdf, precision=16):
    from oracle import coder
    cdf = np.asarray(cdf)
    rows = cdf.reshape(-1, cdf.shape[-1])
    shape = [int(np.asarray(s)) for s in shape]
    n = int(np.prod(shape))
    idx = np.tile(np.arange(rows.shape[0], dtype=np.int32), n // rows.shape[0]) if cdf.shape[0] == 1 else np.arange(n, dtype=np.int32)
    return coder.range_decode(bytes(strings), n, rows, idx, precision).reshape(shape)


def install():
    """Register the shim as ``tensorflow`` (+ the contrib.coder import path) in sys.modules."""
    tf = types.ModuleType("tensorflow")
    tf.float32, tf.int32, tf.int16, tf.string = np.float32, np.int32, np.int16, object
    tf.TensorShape = TensorShape
    tf.Graph = lambda: None
    tf.enable_eager_execution = lambda: None

    keras = types.ModuleType("tensorflow.keras")
    layers = types.ModuleType("tensorflow.keras.layers")
    layers.Layer, layers.InputSpec, layers.Conv3D, layers.Conv3DTranspose = Layer, InputSpec, Conv3D, Conv3DTranspose
    keras.layers, keras.Model = layers, Model
    tf.keras = keras

    init = types.ModuleType("tensorflow.initializers")
    init.constant = lambda v: _Initializer("constant", v)
    init.random_uniform = lambda a, b: _Initializer("uniform", a, b)
    init.zeros = lambda: _Initializer("zeros")
    tf.initializers = init

    nn = types.ModuleType("tensorflow.nn")
    nn.softplus = lambda x: np.logaddexp(_np(x), np.zeros_like(_np(x)))
    nn.relu = lambda x: np.maximum(_np(x), 0)
    tf.nn = nn

    linalg = types.ModuleType("tensorflow.linalg")
    linalg.matmul = lambda a, b: np.matmul(a, b)
    tf.linalg = linalg

    m = types.ModuleType("tensorflow.math")
    m.tanh, m.sign, m.abs, m.exp = np.tanh, np.sign, np.abs, np.exp
    m.sigmoid = _sigmoid
    m.round = np.rint                      # tf.math.round: half to even
    m.add_n = lambda xs: sum(xs[1:], xs[0])
    m.greater = lambda a, b: _np(a) > _np(b)
    m.less_equal = lambda a, b: _np(a) <= _np(b)
    tf.math = m

    rnd = types.ModuleType("tensorflow.random")
    rnd.uniform = lambda shape, a, b: np.random.uniform(a, b, shape).astype(np.float32)
    tf.random = rnd

    tf.abs, tf.exp = np.abs, np.exp
    tf.transpose = lambda x, perm: np.transpose(x, perm)
    tf.reshape = lambda x, shape: np.reshape(x, [int(s) for s in shape])
    tf.shape = lambda x: np.array(np.shape(x), np.int32)
    tf.constant = lambda v, dtype=None: np.asarray(v, dtype=dtype)[()] if dtype else np.asarray(v)[()]
    tf.convert_to_tensor = lambda v, dtype=None: np.asarray(v, dtype=None if dtype in (None, object) else dtype)
    tf.maximum = np.maximum
    tf.range = lambda a, b: np.arange(int(a), int(b), dtype=np.int32)
    tf.tile = lambda x, reps: np.tile(x, [int(r) for r in reps])
    tf.cast = lambda x, dtype: _np(x).astype(dtype)
    tf.reduce_min, tf.reduce_max = np.min, np.max
    tf.reduce_prod = np.prod
    tf.floor, tf.ceil = np.floor, np.ceil
    tf.expand_dims = np.expand_dims
    tf.concat = lambda xs, axis: np.concatenate(xs, axis=axis)
    tf.zeros = lambda shape: np.zeros(shape, np.float32)
    tf.name_scope = _NameScope

    coder_ops = types.ModuleType("tensorflow.contrib.coder.python.ops.coder_ops")
    coder_ops.pmf_to_quantized_cdf = _pmf_to_quantized_cdf
    coder_ops.range_encode = _range_encode
    coder_ops.range_decode = _range_decode
    chain = ["tensorflow.contrib", "tensorflow.contrib.coder", "tensorflow.contrib.coder.python",
             "tensorflow.contrib.coder.python.ops"]
    sys.modules["tensorflow"] = tf
    parent = tf
    for name in chain:
        mod = types.ModuleType(name)
        setattr(parent, name.rsplit(".", 1)[-1], mod)
        sys.modules[name] = mod
        parent = mod
    parent.coder_ops = coder_ops
    sys.modules["tensorflow.contrib.coder.python.ops.coder_ops"] = coder_ops
    for mod in (keras, layers, init, nn, linalg, m, rnd):
        sys.modules[mod.__name__] = mod
    return tf

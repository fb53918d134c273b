"""
A NumPy stand-in for the ~25 TensorFlow / Keras calls that the reference's model definition makes
(/root/reference/genomad/neural_network/{model,igloo}.py), so that THE REFERENCE'S OWN SOURCE can be executed in a container that
has neither TensorFlow nor Keras (tests/golden/make_reference_graph_golden.py).  Test infrastructure only.

What this pins and what it does not: the layer sequence, shapes, weight creation order, patch gather indexing, transposes,
reshapes, softmax axis ... come from the reference's code, not from a restatement; the semantics of the library calls themselves
(below, each with the Keras / TF definition it follows) are this file's.  TensorFlow's own fp32 arithmetic stays unpinned.

Semantics implemented (Keras 3 / TF 2 documentation):
  tf.one_hot(x, depth, axis=-1) float32;  tf.matmul = numpy matmul with batch broadcasting;  tf.transpose(x, perm);
  tf.gather_nd(params, indices): indices [..., 1] -> params[indices[..., 0]] (index depth 1 gathers along axis 0);
  tf.multiply / reshape / squeeze / expand_dims;  tf.nn.softmax over the last axis (max-subtracted);
  Conv1D(filters, k, padding="causal"): left-pad k-1 zeros, cross-correlation y[t] = b + sum_j x[t-(k-1)+j] @ W[j], W [k][in][out];
  LeakyReLU(negative_slope);  Dropout / SpatialDropout1D: identity at inference;  MaxPool1D(pool): stride = pool, padding "valid";
  Concatenate(): last axis;  Dense(units, activation): x @ kernel + bias;  BatchNormalization(): inference form with
  epsilon = 1e-3 (Keras default), (x - moving_mean) / sqrt(moving_variance + eps) * gamma + beta;  Activation("relu"|"softmax").
  Layer.add_weight / build / call, layer.weights = trainable then non-trainable variables (creation order);
  Model(inputs, outputs): functional graph, .layers in creation order, nested models are layers;
  Model.load_weights(h5): the legacy Keras-2 H5 layout; saved names `<layer>/<variable>:0` are paired with the model's layers by
  their default Keras names (a function of creation order in the reference's code) and variables by name, every variable exactly
  once with the saved shape (the reader is genomad_b200/h5lite.py).
"""
import re
import sys
import types

import numpy as np

FLOAT = np.float32          # arithmetic type of the stand-in; set_float(np.float64) gives the same graph in double precision


def set_float(dtype):
    global FLOAT
    FLOAT = dtype


# ------------------------------------------------------------------------------------------ tensors of the functional graph
class Sym:
    """A symbolic tensor: the node that produces it and a sample value (batch 1) for shape inference."""

    def __init__(self, layer, inputs, sample):
        self.layer, self.inputs, self.sample = layer, inputs, sample

    @property
    def shape(self):
        return (None,) + tuple(self.sample.shape[1:])


def _is_sym(x):
    return isinstance(x, Sym) or (isinstance(x, (list, tuple)) and len(x) > 0 and all(isinstance(v, Sym) for v in x))


class Variable:
    def __init__(self, value, name, trainable):
        self.value, self.name, self.trainable = value, name, trainable

    @property
    def shape(self):
        return self.value.shape

    def __array__(self, dtype=None, copy=None):
        return self.value if dtype is None else self.value.astype(dtype)


def _val(x):
    v = x.value if isinstance(x, Variable) else np.asarray(x)
    return v.astype(FLOAT) if v.dtype.kind == "f" and v.dtype != FLOAT else v


_counters = {}


def _auto_name(cls_name):
    """Keras' default layer names: snake-cased class name + a per-name counter (conv1d, conv1d_1, igloo1d_kernel, dense_2 ... --
    the names the shipped file was saved under, which is what load_weights below matches on)."""
    import re
    snake = re.sub(r"([a-z])([A-Z])", r"\1_\2", re.sub(r"(.)([A-Z][a-z]+)", r"\1_\2", cls_name)).lower()
    n = _counters.get(snake, 0)
    _counters[snake] = n + 1
    return snake if n == 0 else f"{snake}_{n}"


_created = []          # every layer in creation order (Model.layers is filtered from this)


class Layer:
    def __init__(self, name=None, **kw):
        self.name = name or _auto_name(type(self).__name__)
        self.trainable = True
        self.built = False
        self._vars = []
        _created.append(self)

    # -- Keras API used by the reference
    def add_weight(self, shape=None, initializer=None, trainable=True, regularizer=None, name=None, dtype=None):
        shape = tuple(int(s) for s in shape)
        if callable(initializer):
            value = np.asarray(initializer(shape, dtype=dtype))
            assert value.shape == shape, (name, value.shape, shape)
            value = value.astype(dtype or np.float32)
        else:
            value = np.zeros(shape, dtype or np.float32)          # placeholders: every weight is loaded from the file afterwards
        v = Variable(value, name, trainable)
        self._vars.append(v)
        return v

    @property
    def weights(self):
        return [v for v in self._vars if v.trainable] + [v for v in self._vars if not v.trainable]

    def build(self, input_shape):
        pass

    def call(self, x):
        raise NotImplementedError

    def _shape_of(self, x):
        if isinstance(x, (list, tuple)):
            return [self._shape_of(v) for v in x]
        return (None,) + tuple(np.asarray(x).shape[1:])

    def _run(self, x):
        if not self.built:
            self.build(self._shape_of(x))
            self.built = True
        return self.call(x)

    def __call__(self, x):
        if _is_sym(x):
            sample = [v.sample for v in x] if isinstance(x, (list, tuple)) else x.sample
            return Sym(self, x, self._run(sample))
        return self._run(x)                                   # eager use inside another layer's call()


class InputLayer(Layer):
    pass


def Input(shape=None, dtype=None, name=None):
    lay = InputLayer(name=name)
    sample = np.zeros((1,) + tuple(shape), dtype=np.int64 if dtype in ("int64", np.int64) else np.float32)
    return Sym(lay, None, sample)


class Conv1D(Layer):
    def __init__(self, filters, kernel_size, padding="valid", **kw):
        super().__init__(**kw)
        assert padding == "causal"
        self.filters, self.k = int(filters), int(kernel_size)

    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(self.k, input_shape[-1], self.filters), name="kernel")
        self.bias = self.add_weight(shape=(self.filters,), name="bias")

    def call(self, x):
        x = np.asarray(x, FLOAT)
        b, t, c = x.shape
        xp = np.concatenate([np.zeros((b, self.k - 1, c), FLOAT), x], axis=1)
        y = np.zeros((b, t, self.filters), FLOAT) + _val(self.bias)
        w = _val(self.kernel)
        for j in range(self.k):
            y += xp[:, j:j + t, :] @ w[j]
        return y


class LeakyReLU(Layer):
    def __init__(self, negative_slope=0.3, **kw):
        super().__init__(**kw)
        self.slope = negative_slope

    def call(self, x):
        x = np.asarray(x, FLOAT)
        return np.where(x > 0, x, x * FLOAT(self.slope)).astype(FLOAT)


class _Identity(Layer):
    def __init__(self, rate=None, **kw):
        super().__init__(**kw)

    def call(self, x):
        return x


class Dropout(_Identity):
    pass


class SpatialDropout1D(_Identity):
    pass


class MaxPool1D(Layer):
    def __init__(self, pool_size=2, **kw):
        super().__init__(**kw)
        self.pool = int(pool_size)

    def call(self, x):
        x = np.asarray(x)
        b, t, c = x.shape
        n = t // self.pool
        return x[:, :n * self.pool].reshape(b, n, self.pool, c).max(axis=2)


MaxPooling1D = MaxPool1D


class Concatenate(Layer):
    def call(self, xs):
        return np.concatenate([np.asarray(v) for v in xs], axis=-1)


def _softmax(x):
    x = np.asarray(x, FLOAT)
    e = np.exp(x - x.max(axis=-1, keepdims=True))
    return (e / e.sum(axis=-1, keepdims=True)).astype(FLOAT)


def _activation(name):
    if name is None or name == "linear":
        return lambda v: v
    if name == "relu":
        return lambda v: np.maximum(v, FLOAT(0))
    if name == "softmax":
        return _softmax
    raise NotImplementedError(name)


class Dense(Layer):
    def __init__(self, units, activation=None, **kw):
        super().__init__(**kw)
        self.units, self.act = int(units), _activation(activation)

    def build(self, input_shape):
        self.kernel = self.add_weight(shape=(input_shape[-1], self.units), name="kernel")
        self.bias = self.add_weight(shape=(self.units,), name="bias")

    def call(self, x):
        return self.act((np.asarray(x, FLOAT) @ _val(self.kernel) + _val(self.bias)).astype(FLOAT))


class BatchNormalization(Layer):
    def __init__(self, epsilon=1e-3, **kw):
        super().__init__(**kw)
        self.eps = epsilon

    def build(self, input_shape):
        d = input_shape[-1]
        self.gamma = self.add_weight(shape=(d,), name="gamma")
        self.beta = self.add_weight(shape=(d,), name="beta")
        self.moving_mean = self.add_weight(shape=(d,), name="moving_mean", trainable=False)
        self.moving_variance = self.add_weight(shape=(d,), name="moving_variance", trainable=False)

    def call(self, x):
        inv = _val(self.gamma) / np.sqrt(_val(self.moving_variance) + FLOAT(self.eps))
        return ((np.asarray(x, FLOAT) - _val(self.moving_mean)) * inv + _val(self.beta)).astype(FLOAT)


class Activation(Layer):
    def __init__(self, activation, **kw):
        super().__init__(**kw)
        self.act = _activation(activation)

    def call(self, x):
        return self.act(np.asarray(x, FLOAT))


class Model(Layer):
    """Functional model.  .layers = the layers on the path from inputs to outputs, in creation order."""

    def __init__(self, inputs=None, outputs=None, **kw):
        super().__init__(**kw)
        self.inputs, self.outputs = inputs, outputs
        on_path = set()

        def walk(s):
            if id(s.layer) in on_path and s.inputs is None:
                return
            on_path.add(id(s.layer))
            if s.inputs is not None:
                for v in (s.inputs if isinstance(s.inputs, (list, tuple)) else [s.inputs]):
                    walk(v)
        walk(outputs)
        self.layers = [l for l in _created if id(l) in on_path and l is not self]
        self.built = True

    @property
    def weights(self):
        return [w for l in self.layers for w in l.weights]

    def _eval(self, s, feed, cache):
        if id(s) in cache:
            return cache[id(s)]
        if s.inputs is None:
            out = feed
        elif isinstance(s.inputs, (list, tuple)):
            out = s.layer._run([self._eval(v, feed, cache) for v in s.inputs])
        else:
            out = s.layer._run(self._eval(s.inputs, feed, cache))
        cache[id(s)] = out
        return out

    def call(self, x):
        return self._eval(self.outputs, x, {})

    def predict(self, x, batch_size=32, verbose=0):
        x = np.asarray(x)
        return np.concatenate([self.call(x[i:i + batch_size]) for i in range(0, len(x), batch_size)])

    # Legacy Keras-2 H5 (`layer_names` / `weight_names` attributes, datasets /<layer>/<saved name>).  Keras' own legacy loader pairs
    # file and model BY ORDER (model.layers sorted by graph depth -- the file lists conv1d, conv1d_1, conv1d_2 before the two IGLOO
    # kernels); saved names are `<layer name>/<variable>:0`, and default layer names are a pure function of the creation order in
    # the reference's code, so pairing BY NAME is the same assignment without having to re-implement the depth sort.  Every
    # variable of the model must be hit exactly once, with the saved shape.
    def load_weights(self, path):
        from genomad_b200.h5lite import H5File
        f = H5File(path)

        def strs(v):
            return [x.decode() if isinstance(x, bytes) else str(x) for x in v]

        def all_layers(m):
            out = []
            for l in m.layers:
                out.append(l)
                if isinstance(l, Model):
                    out.extend(all_layers(l))
            return out
        # canonical names = the default names this model tree would have as the first model of a process (a second
        # create_classifier() in the same process gets conv1d_3 ...; Keras' by-order loader does not care, so neither may this one)
        by_name, per_base = {}, {}
        order = {id(l): i for i, l in enumerate(_created)}
        for l in sorted(all_layers(self), key=lambda l: order[id(l)]):
            base = re.sub(r"_\d+$", "", l.name) if re.sub(r"_\d+$", "", l.name) in _counters else l.name
            k = per_base.get(base, 0)
            per_base[base] = k + 1
            canon = base if k == 0 else f"{base}_{k}"
            assert canon not in by_name, canon
            by_name[canon] = l
        self.load_report = []
        hit = set()
        for ln in strs(f.attrs["/"]["layer_names"]):
            for n in strs(f.attrs.get("/" + ln, {}).get("weight_names", [])):
                arr = f.datasets[f"/{ln}/{n}"]
                lname, vname = n.split("/")[0], n.split("/")[1].split(":")[0]
                layer = by_name[lname]
                var = [v for v in layer._vars if v.name == vname]
                if len(var) != 1:
                    raise ValueError(f"/{ln}/{n}: layer {lname} has no single variable {vname}")
                v = var[0]
                if tuple(v.shape) != tuple(arr.shape):
                    raise ValueError(f"/{ln}/{n}: saved shape {arr.shape} vs {lname}/{vname} {v.shape}")
                if id(v) in hit:
                    raise ValueError(f"/{ln}/{n}: variable assigned twice")
                hit.add(id(v))
                v.value = np.ascontiguousarray(arr)
                self.load_report.append((f"/{ln}/{n}", f"{lname}/{vname}", tuple(arr.shape), str(arr.dtype)))
        missing = [f"{l.name}/{v.name}" for l in all_layers(self) if not isinstance(l, Model) for v in l._vars if id(v) not in hit]
        if missing:
            raise ValueError(f"variables not in the file: {missing}")


# ------------------------------------------------------------------------------------------ the `tensorflow` functions
def _np(x):
    return _val(x)


def _tf_module():
    tf = types.ModuleType("tensorflow")

    def one_hot(x, depth, axis=-1):
        assert axis == -1
        x = np.asarray(_np(x)).astype(np.int64)
        return (x[..., None] == np.arange(depth)).astype(FLOAT)
    tf.one_hot = one_hot
    tf.matmul = lambda a, b: np.matmul(_np(a), _np(b))
    tf.transpose = lambda a, perm=None: np.transpose(_np(a), perm)

    def gather_nd(params, indices):
        idx = np.asarray(_np(indices))
        assert idx.shape[-1] == 1                              # index depth 1: gather whole slices along axis 0
        return _np(params)[idx[..., 0]]
    tf.gather_nd = gather_nd
    tf.multiply = lambda a, b: _np(a) * _np(b)
    tf.reshape = lambda a, shape: np.reshape(_np(a), shape)
    tf.squeeze = lambda a, axis=None: np.squeeze(_np(a), axis=axis)
    tf.expand_dims = lambda a, axis: np.expand_dims(_np(a), axis)
    tf.reduce_mean = lambda a, axis=None: np.mean(np.asarray([_np(v) for v in a]) if isinstance(a, list) else _np(a), axis=axis)
    tf.nn = types.SimpleNamespace(softmax=lambda a: _softmax(_np(a)))
    tf.float32, tf.int64 = "float32", "int64"
    _add_module_driver_calls(tf)
    return tf


def _add_module_driver_calls(tf):
    """The calls genomad/modules/nn_classification.py makes around the model: thread settings (no-ops), the TFRecord writer and
    tf.train.Example (bytes via oracle/tfrecord.py, which tests/test_tfrecord_cpu.py checks against the real protobuf runtime),
    FixedLenFeature / parse_single_example, a TFRecordDataset with map / batch / prefetch, gfile.glob, concat and segment_mean
    (per-segment fp32 sum divided by the count, segment ids sorted as TF requires)."""
    import glob as _glob
    import struct
    from oracle import tfrecord as R
    tf.config = types.SimpleNamespace(threading=types.SimpleNamespace(set_inter_op_parallelism_threads=lambda n: None,
                                                                        set_intra_op_parallelism_threads=lambda n: None))

    class _Int64List:
        def __init__(self, value=()):
            self.value = [int(v) for v in value]

    class _Feature:
        def __init__(self, int64_list=None):
            self.int64_list = int64_list

    class _Features:
        def __init__(self, feature=None):
            self.feature = dict(feature or {})

    class _Example:
        def __init__(self, features=None):
            self.features = features

        def SerializeToString(self):
            assert list(self.features.feature) == ["sequence"]
            return R.serialize_example(self.features.feature["sequence"].int64_list.value)
    tf.train = types.SimpleNamespace(Int64List=_Int64List, Feature=_Feature, Features=_Features, Example=_Example)

    class _Writer:
        def __init__(self, path):
            self.f = open(path, "wb")

        def write(self, record):
            self.f.write(R.frame(record))

        def __enter__(self):
            return self

        def __exit__(self, *a):
            self.f.close()

    class _FixedLenFeature:
        def __init__(self, shape, dtype):
            self.shape, self.dtype = list(shape), dtype

    def _varint(buf, pos):
        v = shift = 0
        while True:
            b = buf[pos]
            pos += 1
            v |= (b & 0x7F) << shift
            shift += 7
            if b < 128:
                return v, pos

    def _unwrap(buf, tag):
        assert buf[0] == tag, (buf[0], tag)
        n, pos = _varint(buf, 1)
        assert pos + n <= len(buf)
        return buf[pos:pos + n], buf[pos + n:]

    def _parse_example(record):
        features, rest = _unwrap(record, 0x0A)                  # Example.features
        assert not rest
        entry, rest = _unwrap(features, 0x0A)                   # Features.feature map entry
        assert not rest
        key, value = _unwrap(entry, 0x0A)
        assert key == b"sequence"
        feature, rest = _unwrap(value, 0x12)
        assert not rest
        int64_list, rest = _unwrap(feature, 0x1A)               # Feature.int64_list
        assert not rest
        packed, rest = _unwrap(int64_list, 0x0A)                # Int64List.value, packed varints
        assert not rest
        out, pos = [], 0
        while pos < len(packed):
            v, pos = _varint(packed, pos)
            out.append(v)
        return out

    def parse_single_example(record, description):
        (name, spec), = description.items()
        tokens = np.asarray(_parse_example(record), np.int64)
        assert name == "sequence" and list(tokens.shape) == spec.shape
        return {name: tokens}

    class _Dataset:
        def __init__(self, it):
            self._it = it

        def __iter__(self):
            return iter(self._it())

        def map(self, fn, num_parallel_calls=None, deterministic=None):
            src = self._it
            return _Dataset(lambda: (fn(x) for x in src()))

        def batch(self, n):
            src = self._it

            def gen():
                buf = []
                for x in src():
                    buf.append(x)
                    if len(buf) == n:
                        yield np.stack(buf)
                        buf = []
                if buf:
                    yield np.stack(buf)
            return _Dataset(gen)

        def prefetch(self, n):
            return self

    def _records(filenames):
        for fn in filenames:
            blob = open(fn, "rb").read()
            pos = 0
            while pos < len(blob):
                (n,) = struct.unpack_from("<Q", blob, pos)
                assert struct.unpack_from("<I", blob, pos + 8)[0] == R.masked_crc(blob[pos:pos + 8])
                data = blob[pos + 12:pos + 12 + n]
                assert struct.unpack_from("<I", blob, pos + 12 + n)[0] == R.masked_crc(data)
                yield data
                pos += 16 + n
    tf.io = types.SimpleNamespace(TFRecordWriter=_Writer, FixedLenFeature=_FixedLenFeature, parse_single_example=parse_single_example,
                                  gfile=types.SimpleNamespace(glob=lambda pat: _glob.glob(pat)))
    tf.data = types.SimpleNamespace(TFRecordDataset=lambda filenames, num_parallel_reads=None: _Dataset(lambda: _records(list(filenames))),
                                    experimental=types.SimpleNamespace(AUTOTUNE=-1))
    tf.concat = lambda values, axis=0: np.concatenate([_np(v) for v in values], axis=axis)

    def segment_mean(data, segment_ids):
        data, ids = np.asarray(_np(data), np.float32), np.asarray(segment_ids).astype(np.int64)
        assert len(ids) == len(data) and np.all(np.diff(ids) >= 0)
        out = np.zeros((int(ids[-1]) + 1,) + data.shape[1:], np.float32)
        cnt = np.zeros(len(out), np.float32)
        for row, i in zip(data, ids):                           # sequential fp32 accumulation
            out[i] += row
            cnt[i] += 1
        return out / np.maximum(cnt, 1)[:, None]
    tf.math = types.SimpleNamespace(segment_mean=segment_mean)


def install():
    """Put the stand-ins into sys.modules as `tensorflow`, `keras`, `keras.layers`, `keras.regularizers`.  Returns the modules."""
    _counters.clear()
    _created.clear()
    tf = _tf_module()
    keras = types.ModuleType("keras")
    layers = types.ModuleType("keras.layers")
    for cls in (Layer, Conv1D, LeakyReLU, Dropout, SpatialDropout1D, MaxPool1D, Concatenate, Dense, BatchNormalization, Activation):
        setattr(layers, cls.__name__, cls)
    layers.MaxPooling1D = MaxPool1D
    layers.Input = Input
    regs = types.ModuleType("keras.regularizers")
    regs.l2 = lambda v: ("l2", v)
    keras.layers, keras.regularizers, keras.Model, keras.Layer = layers, regs, Model, Layer
    sys.modules.update({"tensorflow": tf, "keras": keras, "keras.layers": layers, "keras.regularizers": regs})
    return tf, keras

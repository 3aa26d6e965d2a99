"""TEST INFRASTRUCTURE -- a small stand-in for the slice of TensorFlow 1.x that the reference's MADDPG graph code uses
(maddpg/trainer/maddpg.py:20-110 ``make_update_exp`` / ``p_train`` / ``q_train``, maddpg/common/distributions.py SoftCategorical,
maddpg/common/tf_util.py ``function`` / ``scope_vars`` / ``minimize_and_clip`` / sessions, experiments/train.py:39-46
``mlp_model``), so that THAT CODE -- unmodified -- can be executed in the build container and its results recorded as golden
vectors (tests/golden/make_graph_golden.py).

It is a lazy graph: every ``tf.*`` call returns a ``Tensor`` node; ``Session.run`` evaluates the fetched nodes once per run
(memoised, so a loss and its gradients see the same ``random_uniform`` draw, like a TF run) with torch float32 tensors and torch
autograd for ``compute_gradients``.  What the reference's code decides is therefore executed for real: which tensors are
concatenated into which network, which variables each optimizer owns (``scope_vars`` by name prefix), the loss expressions, where
``clip_by_norm`` sits, how ``make_update_exp`` pairs variables by sorted name, what ``U.function`` feeds and fetches.  What this
file RESTATES, from TensorFlow's documentation, is the primitive semantics: ``fully_connected`` = x @ W + b with variables
``<scope>/fully_connected[_k]/{weights,biases}`` (sub-scope counters reset when the enclosing named scope is left, which is what
makes ``reuse=True`` find them), ``clip_by_norm`` = t * c / max(||t||, c), and ``AdamOptimizer`` = TF's formula with the epsilon
outside the bias correction.  ``random_uniform`` draws from the pluggable ``NOISE`` callback.  Variable
writes of one ``Session.run`` are committed together when the run ends (every gradient of a run is taken at the variables the run
started with).

The second half stands in for the slice of DeepMind Sonnet 1.x the fork's modules use (``snt.AbstractModule``,
``snt.reuse_variables``, ``snt.nets.MLP``) plus the extra TensorFlow calls of maddpg/modules/*.py and maddpg/algorithms/*.py, for
tests/golden/make_fork_graph_golden.py.
"""
import re
import sys
import types

import numpy as np
import torch

NOISE = [lambda shape: np.random.uniform(size=shape).astype(np.float32)]
_VARIABLES = []          # every Variable, creation order
_SCOPES = [("", False)]  # (full name, reuse)
_SUBSCOPE_COUNT = {}     # full scope name -> {default name -> times opened}
_SESSIONS = []
_UNIQUE = {}


def reset():
    del _VARIABLES[:]
    del _SCOPES[1:]
    _SUBSCOPE_COUNT.clear()
    _UNIQUE.clear()


def _unique(name):
    k = _UNIQUE.get(name, 0)
    _UNIQUE[name] = k + 1
    return name if k == 0 else "%s_%d" % (name, k)


class _Ctx(object):
    def __init__(self, feed):
        self.feed, self.cache, self.pending = feed, {}, []   # pending: variable writes, committed when the run ends


class Tensor(object):
    """A graph node.  ``type(x) is tf.Tensor and len(x.op.inputs) == 0`` identifies a placeholder (tf_util.py:58-59)."""

    def __init__(self, fn, inputs=(), name="op", last_dim=None):
        self._fn, self._inputs, self.last_dim = fn, tuple(inputs), last_dim
        full = (_SCOPES[-1][0] + "/" if _SCOPES[-1][0] else "") + name
        self.name = _unique(full) + ":0"
        self.op = types.SimpleNamespace(inputs=self._inputs, name=self.name[:-2])

    def eval(self, ctx):
        key = id(self)
        if key not in ctx.cache:
            ctx.cache[key] = self._fn(ctx, *[i.eval(ctx) if isinstance(i, Tensor) else i for i in self._inputs])
        return ctx.cache[key]

    def _bin(self, other, f, name):
        last = self.last_dim if self.last_dim is not None else getattr(other, "last_dim", None)
        return Tensor(lambda c, a, b: f(a, b), (self, other), name, last)

    def __add__(self, o): return self._bin(o, lambda a, b: a + b, "add")
    def __radd__(self, o): return self._bin(o, lambda a, b: b + a, "add")
    def __sub__(self, o): return self._bin(o, lambda a, b: a - b, "sub")
    def __rsub__(self, o): return self._bin(o, lambda a, b: b - a, "sub")
    def __mul__(self, o): return self._bin(o, lambda a, b: a * b, "mul")
    def __rmul__(self, o): return self._bin(o, lambda a, b: b * a, "mul")
    def __truediv__(self, o): return self._bin(o, lambda a, b: a / b, "div")
    def __neg__(self): return Tensor(lambda c, a: -a, (self,), "neg", self.last_dim)
    def __getitem__(self, idx): return Tensor(lambda c, a: a[idx], (self,), "strided_slice")
    def get_shape(self): return [None, self.last_dim]      # every tensor the reference asks is a (batch, features) matrix
    __hash__ = object.__hash__


class Variable(Tensor):
    def __init__(self, name, shape, init):
        self._fn, self._inputs, self.last_dim = None, (), shape[-1]
        self.name = name + ":0"
        self.op = types.SimpleNamespace(inputs=(), name=name)
        self.value = torch.tensor(np.asarray(init(shape), np.float32), requires_grad=True)
        _VARIABLES.append(self)

    def eval(self, ctx):
        return self.value

    def assign(self, expr, use_locking=None):
        def run(c, v):
            c.pending.append((self, v.detach().clone()))
            return None
        return Tensor(run, (expr,), "Assign")

    def load(self, array):
        with torch.no_grad():
            self.value.copy_(torch.as_tensor(np.asarray(array, np.float32)).view_as(self.value))

    def numpy(self):
        return self.value.detach().numpy().copy()


def _as_node(x):
    return x


# -- scopes ----------------------------------------------------------------------------------------------------------------------
class _VarScopeInfo(object):
    def __init__(self, name):
        self.name = name


class variable_scope(object):
    def __init__(self, name_or_scope, default_name=None, values=None, reuse=None):
        self._name, self._default, self._reuse = name_or_scope, default_name, reuse

    def __enter__(self):
        parent, parent_reuse = _SCOPES[-1]
        name = self._name.name if isinstance(self._name, _VarScopeInfo) else self._name
        if name is None:   # default-named scope: uniquified by how often this name was opened under the parent
            counts = _SUBSCOPE_COUNT.setdefault(parent, {})
            k = counts.get(self._default, 0)
            counts[self._default] = k + 1
            name = self._default if k == 0 else "%s_%d" % (self._default, k)
        full = (parent + "/" if parent else "") + name
        self._full = full
        _SCOPES.append((full, bool(self._reuse) or parent_reuse))
        return _VarScopeInfo(full)

    def __exit__(self, *a):
        full = _SCOPES.pop()[0]
        for k in list(_SUBSCOPE_COUNT):   # leaving a scope closes its sub-scope counters (what makes a later reuse=True line up)
            if k == full or k.startswith(full + "/"):
                del _SUBSCOPE_COUNT[k]
        return False


def get_variable_scope():
    return _VarScopeInfo(_SCOPES[-1][0])


def get_variable(name, shape, initializer):
    full = (_SCOPES[-1][0] + "/" if _SCOPES[-1][0] else "") + name
    if _SCOPES[-1][1] == "auto":     # a Sonnet module's scope: created on the first call, shared by the later ones
        for v in _VARIABLES:
            if v.op.name == full:
                return v
        return Variable(full, list(shape), initializer)
    if _SCOPES[-1][1]:
        for v in _VARIABLES:
            if v.op.name == full:
                return v
        raise ValueError("Variable %s does not exist, or was not created with tf.get_variable()" % full)
    if any(v.op.name == full for v in _VARIABLES):
        raise ValueError("Variable %s already exists, disallowed. Did you mean to set reuse=True?" % full)
    return Variable(full, list(shape), initializer)


def _xavier(shape):
    lim = np.sqrt(6.0 / (shape[0] + shape[1]))
    return np.random.uniform(-lim, lim, size=shape)


def fully_connected(inputs, num_outputs, activation_fn=None, scope=None, reuse=None):
    """tf.contrib.layers.fully_connected defaults: xavier weights, zero biases, variables under <scope>/fully_connected[_k]."""
    with variable_scope(scope, "fully_connected", [inputs], reuse=reuse):
        W = get_variable("weights", [inputs.last_dim, num_outputs], _xavier)
        b = get_variable("biases", [num_outputs], lambda s: np.zeros(s))
    out = Tensor(lambda c, x, w, bb: x @ w + bb, (inputs, W, b), "fully_connected/BiasAdd", num_outputs)
    return activation_fn(out) if activation_fn is not None else out


# -- ops -------------------------------------------------------------------------------------------------------------------------
def placeholder(dtype, shape=None, name=None):
    t = Tensor(None, (), name or "Placeholder", None if not shape else shape[-1])
    t._fn = lambda c: torch.as_tensor(np.asarray(c.feed[t], np.float32))   # fed values are cast to the placeholder's float32
    return t


def concat(values, axis):
    values = list(values)
    return Tensor(lambda c, *xs: torch.cat(xs, dim=axis), values, "concat",
                  sum(v.last_dim for v in values) if axis in (1, -1) else values[0].last_dim)


def reduce_mean(x, axis=None, keep_dims=False):
    return Tensor(lambda c, a: a.mean() if axis is None else a.mean(dim=axis, keepdim=keep_dims), (x,), "Mean")


def square(x):
    return Tensor(lambda c, a: a * a, (x,), "Square", x.last_dim)


def log(x):
    return Tensor(lambda c, a: torch.log(a), (x,), "Log", x.last_dim)


def shape(x):
    return Tensor(lambda c, a: tuple(a.shape), (x,), "Shape")


def random_uniform(shp):
    return Tensor(lambda c, s: torch.as_tensor(np.asarray(NOISE[0](tuple(s)), np.float32)), (shp,), "random_uniform")


def clip_by_norm(t, clip_norm):
    def run(c, a):
        n = torch.sqrt(torch.sum(a * a))
        return a * clip_norm / torch.maximum(n, torch.tensor(clip_norm, dtype=a.dtype))
    return Tensor(run, (t,), "clip_by_norm")


def group(*ops, name=None):
    flat = []
    for o in ops:
        flat += list(o) if isinstance(o, (list, tuple)) else [o]
    return Tensor(lambda c, *a: None, flat or (1.0,), "group_deps")


def global_variables():
    return list(_VARIABLES)


def variables_initializer(variables):
    return Tensor(lambda c: None, (), "init")


def get_collection(key, scope=None):
    return [v for v in _VARIABLES if scope is None or re.match(scope, v.name)]


class _Gradients(object):
    """One backward pass per run for all of an optimizer's variables."""

    def __init__(self, loss, var_list):
        self.loss, self.var_list = loss, var_list

    def node(self, i):
        def run(c):
            key = ("grads", id(self))
            if key not in c.cache:
                c.cache[key] = torch.autograd.grad(self.loss.eval(c), [v.value for v in self.var_list], retain_graph=True)
            return c.cache[key][i]
        # the loss is listed as an input so that the node is not mistaken for a placeholder
        return Tensor(lambda c, _loss: run(c), (self.loss,), "gradients")


class AdamOptimizer(object):
    def __init__(self, learning_rate=0.001, beta1=0.9, beta2=0.999, epsilon=1e-8, use_locking=False):
        self.lr, self.b1, self.b2, self.eps, self.t, self.slots = learning_rate, beta1, beta2, epsilon, 0, {}

    def compute_gradients(self, loss, var_list=None):
        g = _Gradients(loss, var_list)
        return [(g.node(i), v) for i, v in enumerate(var_list)]

    def apply_gradients(self, grads_and_vars):
        grads_and_vars = list(grads_and_vars)

        def run(c, *grads):
            self.t += 1
            lr_t = np.float32(self.lr * np.sqrt(1.0 - self.b2 ** self.t) / (1.0 - self.b1 ** self.t))
            with torch.no_grad():
                for g, (_, v) in zip(grads, grads_and_vars):
                    m, s = self.slots.setdefault(id(v), (torch.zeros_like(v.value), torch.zeros_like(v.value)))
                    m.mul_(np.float32(self.b1)).add_(np.float32(1.0 - self.b1) * g)
                    s.mul_(np.float32(self.b2)).add_(np.float32(1.0 - self.b2) * g * g)
                    c.pending.append((v, v.value.detach() - lr_t * m / (torch.sqrt(s) + np.float32(self.eps))))
            return None
        return Tensor(run, [g for g, _ in grads_and_vars], "Adam")



NORMAL = [lambda shape: np.random.standard_normal(size=shape).astype(np.float32)]   # N(0, 1) draws behind tf.random.normal


def tanh(x):
    return Tensor(lambda c, a: torch.tanh(a), (x,), "Tanh", x.last_dim)


def clip_by_value(x, lo, hi):
    return Tensor(lambda c, a: torch.clamp(a, lo, hi), (x,), "clip_by_value", x.last_dim)


def random_normal(shp, mean=0.0, stddev=1.0):
    return Tensor(lambda c, s: torch.as_tensor(np.asarray(NORMAL[0](tuple(s)), np.float32)) * np.float32(stddev) + np.float32(mean),
                  (shp,), "random_normal")


def stack(values, axis=0):
    return Tensor(lambda c, *xs: torch.stack(xs, dim=axis), values, "stack")


def reduce_min(x, axis=None):
    return Tensor(lambda c, a: a.min() if axis is None else a.min(dim=axis).values, (x,), "Min")


def reduce_std(x, axis=None):
    return Tensor(lambda c, a: a.std(unbiased=False) if axis is None else a.std(dim=axis, unbiased=False), (x,), "reduce_std")


def stop_gradient(x):
    return Tensor(lambda c, a: a.detach(), (x,), "StopGradient", x.last_dim)


def split(value, num, axis=0):
    if np.ndim(num) == 0:
        return [Tensor((lambda k: lambda c, a: torch.chunk(a, int(num), dim=axis)[k])(k), (value,), "split", value.last_dim)
                for k in range(int(num))]
    sizes = [int(x) for x in num]      # tf.split(value, size_splits, axis)
    return [Tensor((lambda k: lambda c, a: torch.split(a, sizes, dim=axis)[k])(k), (value,), "split", sizes[k]) for k in range(len(sizes))]


def constant(value, dtype=None):
    return Tensor(lambda c, _one: torch.as_tensor(np.asarray(value, np.float32)), (1.0,), "Const")


_CONTROL = [()]


class control_dependencies(object):
    """Nodes created inside run the given ops after their own inputs (the losses a TfFunction returns are computed from the
    pre-step variables: the same forward pass feeds the gradients)."""

    def __init__(self, ops):
        self.ops = tuple(ops)

    def __enter__(self):
        _CONTROL.append(_CONTROL[-1] + self.ops)

    def __exit__(self, *a):
        _CONTROL.pop()
        return False


def identity(x):
    deps = _CONTROL[-1]

    def run(c, a):
        for d in deps:
            d.eval(c)
        return a
    return Tensor(run, (x,), "Identity", getattr(x, "last_dim", None))


def no_op():
    deps = _CONTROL[-1]

    def run(c, _one):
        for d in deps:
            d.eval(c)
        return None
    return Tensor(run, (1.0,), "NoOp")


class Graph(object):
    def as_default(self):
        return self

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def global_variables_initializer():
    return Tensor(lambda c, _one: None, (1.0,), "init")


class Session(object):
    def __init__(self, config=None, graph=None):
        self.graph = graph or Graph()

    def __enter__(self):
        _SESSIONS.append(self)
        return self

    def __exit__(self, *a):
        _SESSIONS.pop()
        return False

    def run(self, fetches, feed_dict=None):
        ctx = _Ctx(feed_dict or {})
        if isinstance(fetches, dict):
            keys = list(fetches)
            return dict(zip(keys, self.run([fetches[k] for k in keys], feed_dict)))
        single = not isinstance(fetches, (list, tuple))
        out = []
        for f in ([fetches] if single else fetches):   # in fetch order: outputs before the update group (tf_util.py:325-326)
            r = f.eval(ctx)
            out.append(r.detach().numpy().copy() if torch.is_tensor(r) else r)
        with torch.no_grad():   # every read of a run sees the variables as they were when it started; writes land together
            for var, value in ctx.pending:
                var.value.copy_(value)
        return out[0] if single else out


def get_default_session():
    return _SESSIONS[-1]


def install():
    """Registers the stand-in as ``tensorflow`` (+ the sub-modules the reference imports) and returns it."""
    tf = types.ModuleType("tensorflow")
    for k, v in globals().items():
        if not k.startswith("_") and k not in ("re", "sys", "types", "np", "torch", "install", "install_sonnet"):
            setattr(tf, k, v)
    tf.float32, tf.int32 = "float32", "int32"
    tf.nn = types.SimpleNamespace(relu=lambda x: Tensor(lambda c, a: torch.relu(a), (x,), "Relu", x.last_dim),
                                  softmax=lambda x, axis=None: Tensor(lambda c, a: torch.softmax(a, dim=-1 if axis is None else axis),
                                                                      (x,), "Softmax", x.last_dim))
    tf.train = types.SimpleNamespace(AdamOptimizer=AdamOptimizer, Saver=lambda *a, **k: object())
    tf.GraphKeys = types.SimpleNamespace(GLOBAL_VARIABLES="variables", TRAINABLE_VARIABLES="trainable_variables")
    tf.ConfigProto = lambda **k: types.SimpleNamespace(gpu_options=types.SimpleNamespace())
    tf.random = types.SimpleNamespace(normal=random_normal)
    tf.math = types.SimpleNamespace(reduce_std=reduce_std)
    contrib, layers = types.ModuleType("tensorflow.contrib"), types.ModuleType("tensorflow.contrib.layers")
    layers.fully_connected = fully_connected
    contrib.layers = layers
    tf.contrib = contrib
    python, ops = types.ModuleType("tensorflow.python"), types.ModuleType("tensorflow.python.ops")
    ops.math_ops, ops.nn = types.SimpleNamespace(), types.SimpleNamespace()
    python.ops = ops
    tf.python = python
    sys.modules.update({"tensorflow": tf, "tensorflow.contrib": contrib, "tensorflow.contrib.layers": layers,
                        "tensorflow.python": python, "tensorflow.python.ops": ops})
    return tf


# -- the slice of DeepMind Sonnet 1.x the fork's modules use -----------------------------------------------------------------------
class AbstractModule(object):
    """snt.AbstractModule: the module's variable scope is fixed at construction (uniquified default name); every call to the
    module -- and every ``snt.reuse_variables`` method -- runs inside that scope and shares its variables."""

    def __init__(self, _sentinel=None, custom_getter=None, name=None):
        with variable_scope(None, default_name=name or type(self).__name__.lower()) as scope:
            self._scope_name = scope.name
        self._connected = False

    def _enter(self):
        return _ModuleScope(self._scope_name)

    def __call__(self, *args, **kwargs):
        with self._enter():
            out = self._build(*args, **kwargs)
        self._connected = True
        return out


class _ModuleScope(object):
    """Re-enters a module's absolute scope; variables are created on first use and found afterwards (template semantics)."""

    def __init__(self, full):
        self.full = full

    def __enter__(self):
        _SCOPES.append((self.full, "auto"))

    def __exit__(self, *a):
        _SCOPES.pop()
        return False


def reuse_variables(method):
    def wrapped(self, *args, **kwargs):
        with self._enter():
            return method(self, *args, **kwargs)
    return wrapped


class Linear(AbstractModule):
    def __init__(self, output_size, name="linear"):
        super().__init__(name=name)
        self.output_size = output_size

    def _build(self, inputs):
        self.w = get_variable("w", [inputs.last_dim, self.output_size], lambda s: np.random.standard_normal(s) / np.sqrt(s[0]))
        self.b = get_variable("b", [self.output_size], lambda s: np.zeros(s))
        return Tensor(lambda c, x, w, b: x @ w + b, (inputs, self.w, self.b), "linear/add", self.output_size)


class MLP(AbstractModule):
    """snt.nets.MLP(output_sizes): Linear layers linear_0..linear_{n-1}, ReLU between them, no final activation."""

    def __init__(self, output_sizes, name="mlp"):
        super().__init__(name=name)
        with self._enter():
            self._layers = [Linear(n, name="linear_%d" % i) for i, n in enumerate(output_sizes)]

    def _build(self, inputs):
        net = inputs
        for i, layer in enumerate(self._layers):
            net = layer(net)
            if i + 1 < len(self._layers):
                net = Tensor(lambda c, a: torch.relu(a), (net,), "Relu", net.last_dim)
        return net

    @property
    def trainable_variables(self):
        out = []
        for layer in self._layers:
            out += [layer.w, layer.b]
        return tuple(out)


def install_sonnet():
    snt = types.ModuleType("sonnet")
    snt.AbstractModule, snt.reuse_variables, snt.Linear = AbstractModule, reuse_variables, Linear
    snt.nets = types.SimpleNamespace(MLP=MLP)
    snt.BatchNormV2 = None
    sys.modules["sonnet"] = snt
    return snt

"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

A minimal stand-in for the `tensorflow` module (TF 1.15 is the reference's only
compute dependency, README.md:18,23, and cannot be installed here: no network,
no cp312 wheel).  Installing it in `sys.modules` lets the reference's own
`ops.py` be imported *unchanged* from /root/reference and its pure
slicing/concat/stack functions (`curl` ops.py:264-274, `jacobian` :205-225,
`jacobian3` :227-262, `divergence` :276-284, `divergence3` :286-290, `lrelu`
:9-10, `upscale`/`upscale3` :66-91) be executed verbatim on torch CPU tensors,
forward and (through torch autograd) backward.

Used only by `oracle/make_golden.py` (in the build container, where
/root/reference exists) to pin `oracle/ref_ops.py` and to write the committed
fixtures under `tests/golden/`.  Everything the reference delegates to TF
*library kernels* (slim.conv2d/conv3d/fully_connected, AdamOptimizer) is NOT
covered by this shim -- see `oracle/ref_ops.py` for the restatement and the
"parity unpinned" note there.  `install_structural` (used by
`oracle/make_golden_model.py`) additionally lets the reference's `model.py` run
unchanged with the layers' arithmetic delegated to the restated primitives: it
pins the model STRUCTURE and variable layout, not the primitives.
"""
import importlib.machinery
import sys
import types

import numpy as np
import torch


def _t(x):
    return x if isinstance(x, torch.Tensor) else torch.as_tensor(x)


def _mod(name):
    """Fake module WITH a real ModuleSpec: importlib.util.find_spec(name) (torch._dynamo probes 'tensorflow' that way)
    raises ValueError on a sys.modules entry whose __spec__ is None."""
    m = types.ModuleType(name)
    m.__spec__ = importlib.machinery.ModuleSpec(name, None)
    m._dfl_shim = True
    return m


def uninstall():
    """Remove every shim module from sys.modules (tests call this so the fake never outlives them)."""
    for k in [k for k, v in sys.modules.items() if k.split(".")[0] == "tensorflow" and getattr(v, "_dfl_shim", False)]:
        del sys.modules[k]


def install():
    """Install the fake `tensorflow` (+ `tensorflow.contrib.slim`) modules."""
    if "tensorflow" in sys.modules and getattr(sys.modules["tensorflow"], "_dfl_shim", False):
        return sys.modules["tensorflow"]
    tf = _mod("tensorflow")
    tf._dfl_shim = True
    tf.float32 = torch.float32
    tf.concat = lambda values, axis=0, name=None: torch.cat([_t(v) for v in values], dim=axis)
    # .clone(): a TF op returns a NEW tensor; trainer.py:598-608 multiplies y_ in place (`y_ *= ...`, a rebind in TF) after
    # taking expand_dims(y_), which must not alias it
    tf.expand_dims = lambda x, axis=None, name=None: torch.unsqueeze(_t(x), axis).clone()
    tf.stack = lambda values, axis=0, name=None: torch.stack([_t(v) for v in values], dim=axis)
    tf.transpose = lambda x, perm=None: _t(x).permute(*perm)
    tf.maximum = lambda a, b: torch.maximum(_t(a), _t(b))
    tf.reshape = lambda x, shape: _t(x).reshape(*shape)
    tf.abs = torch.abs
    tf.reduce_mean = lambda x, axis=None: torch.mean(x) if axis is None else torch.mean(x, dim=axis)

    image = _mod("tensorflow.image")

    def resize_nearest_neighbor(x, new_size):
        # tf.image.resize_nearest_neighbor, align_corners=False: out[i] = in[floor(i*in/out)]
        x = _t(x)
        b, h, w, c = x.shape
        nh, nw = int(new_size[0]), int(new_size[1])
        iy = torch.div(torch.arange(nh) * h, nh, rounding_mode="floor")
        ix = torch.div(torch.arange(nw) * w, nw, rounding_mode="floor")
        return x[:, iy][:, :, ix]

    image.resize_nearest_neighbor = resize_nearest_neighbor
    tf.image = image

    class _Shape(object):
        def __init__(self, s):
            self._s = list(s)

        def as_list(self):
            return list(self._s)

        ndims = property(lambda self: len(self._s))

    # ops.int_shape (ops.py:96-98) calls tensor.get_shape().as_list()
    if not hasattr(torch.Tensor, "get_shape"):
        torch.Tensor.get_shape = lambda self: _Shape(self.shape)

    nn = _mod("tensorflow.nn")          # model.py:218 names tf.nn.elu in a default argument (NN arch, unused here)
    nn.elu = torch.nn.functional.elu
    tf.nn = nn
    sys.modules["tensorflow.nn"] = nn

    contrib = _mod("tensorflow.contrib")
    slim = _mod("tensorflow.contrib.slim")
    contrib.slim = slim
    tf.contrib = contrib
    sys.modules["tensorflow"] = tf
    sys.modules["tensorflow.image"] = image
    sys.modules["tensorflow.contrib"] = contrib
    sys.modules["tensorflow.contrib.slim"] = slim
    return tf


class VariableStore(object):
    """Variables the reference's model builders ask for, in the order they ask (what tf.get_variable + slim's layer
    scopes would create).  `values` (name -> tensor) must be pre-filled by the caller; a request for a missing name or a
    different shape is an error -- that is how the oracle's layout tables get pinned against the reference's own code."""

    def __init__(self, values):
        self.values = values
        self.requested = []            # [(name, shape)] in creation order
        self.scopes = []

    def full(self, leaf):
        return "/".join(self.scopes + [leaf])

    def get(self, leaf, shape):
        name = self.full(leaf)
        if name not in self.values:
            raise KeyError("the reference asked for variable %r %s which the oracle layout does not have" % (name, tuple(shape)))
        v = self.values[name]
        if tuple(v.shape) != tuple(shape):
            raise ValueError("variable %r: reference shape %s, oracle layout %s" % (name, tuple(shape), tuple(v.shape)))
        if name not in [n for n, _ in self.requested]:
            self.requested.append((name, tuple(shape)))
        return v


def install_structural(store, conv_nd, linear):
    """Extend the shim so that the reference's model.py (GeneratorBE/BE3, EncoderBE/BE3, AE/AE3) runs UNCHANGED:
    tf.variable_scope, slim.conv2d / conv3d / fully_connected (variable creation order, names `<scope>/weights|biases`,
    shapes [k,(k,)k,Cin,Cout] / [in,out]) and tf.contrib.framework.get_variables.  The layers' ARITHMETIC is delegated to
    the oracle's restated primitives `conv_nd` / `linear` (TF's kernels are not available), so this pins the model
    STRUCTURE -- layer order, residual / concat / upsample wiring, strides, variable naming -- not the primitives."""
    import contextlib

    tf = install()
    slim = sys.modules["tensorflow.contrib.slim"]

    class _VS(object):
        def __init__(self, name):
            self.name = name

    # slim layers called with scope=None open variable_scope(None, default_name='Conv' / 'fully_connected'): TF makes the
    # name unique WITHIN the enclosing variable scope (Conv, Conv_1, ...) and forgets those counts when the enclosing scope
    # closes, which is what lets DiscriminatorPatch(..., reuse=True) (model.py:89-116) find D/Conv ... D/Conv_4 again.
    default_counts = {}

    def _unique_default(default_name):
        cnt = default_counts.setdefault("/".join(store.scopes), {})
        i = cnt.get(default_name, 0)
        cnt[default_name] = i + 1
        return default_name if i == 0 else "%s_%d" % (default_name, i)

    @contextlib.contextmanager
    def variable_scope(name, reuse=None):
        store.scopes.append(name)
        path = "/".join(store.scopes)
        try:
            yield _VS(path)
        finally:
            for k in [k for k in default_counts if k == path or k.startswith(path + "/")]:
                del default_counts[k]
            store.scopes.pop()

    def _conv(nd):
        def conv(x, o_dim, k, stride=1, activation_fn=None, scope=None, data_format=None):
            x = _t(x)
            assert data_format in (None, "NHWC", "NDHWC"), data_format
            store.scopes.append(scope if scope is not None else _unique_default("Conv"))
            try:
                w = store.get("weights", (k,) * nd + (x.shape[-1], o_dim))
                b = store.get("biases", (o_dim,))
            finally:
                store.scopes.pop()
            return conv_nd(x, w, b, stride, activation_fn)
        return conv

    def fully_connected(x, o_dim, activation_fn=None, scope=None):
        x = _t(x)
        store.scopes.append(scope if scope is not None else _unique_default("fully_connected"))
        try:
            w = store.get("weights", (x.shape[-1], o_dim))
            b = store.get("biases", (o_dim,))
        finally:
            store.scopes.pop()
        y = linear(x, w, b)
        return activation_fn(y) if activation_fn is not None else y

    slim.conv2d, slim.conv3d, slim.fully_connected = _conv(2), _conv(3), fully_connected
    tf._dfl_unique_default = _unique_default
    tf.variable_scope = variable_scope
    tf.sigmoid = torch.sigmoid
    framework = _mod("tensorflow.contrib.framework")
    framework.get_variables = lambda vs: [n for n, _ in store.requested if n == vs.name or n.startswith(vs.name + "/")]
    sys.modules["tensorflow"].contrib.framework = framework
    sys.modules["tensorflow.contrib.framework"] = framework
    return tf


def install_nn(store, batch_norm, dropout, masks, mse):
    """Shim surface of arch=nn (model.py:218-224, trainer.py:586-629): slim.batch_norm (variables beta, gamma,
    moving_mean, moving_variance under the default scope BatchNorm / BatchNorm_1), slim.dropout, tf.add,
    tf.losses.mean_squared_error.  Arithmetic is delegated to the oracle's restated primitives; `masks` is a list the
    dropout masks are popped from in call order (training-mode calls only).  Call install_structural() first."""
    tf = sys.modules["tensorflow"]
    slim = sys.modules["tensorflow.contrib.slim"]
    _default = tf._dfl_unique_default      # per enclosing scope, forgotten when it closes (NN(..., reuse=True) starts over)

    def slim_batch_norm(x, decay=0.999, updates_collections="update_ops", epsilon=0.001, scale=False, fused=None,
                        is_training=True, activation_fn=None, data_format="NHWC", scope=None):
        assert updates_collections is None and scale and fused and data_format == "NHWC"
        x = _t(x)
        store.scopes.append(scope if scope is not None else _default("BatchNorm"))
        try:
            n = x.shape[-1]
            beta, gamma = store.get("beta", (n,)), store.get("gamma", (n,))
            mm, mv = store.get("moving_mean", (n,)), store.get("moving_variance", (n,))
        finally:
            store.scopes.pop()
        return batch_norm(x, gamma, beta, mm, mv, is_training, epsilon, decay, activation_fn)

    def slim_dropout(x, keep_prob=0.5, is_training=True):
        return dropout(_t(x), keep_prob, masks.pop(0)) if is_training else _t(x)

    slim.batch_norm, slim.dropout = slim_batch_norm, slim_dropout
    tf.add = lambda a, b: _t(a) + _t(b)
    losses = _mod("tensorflow.losses")
    losses.mean_squared_error = lambda labels, predictions: mse(_t(labels), _t(predictions))
    tf.losses = losses
    sys.modules["tensorflow.losses"] = losses


class BuildDone(Exception):
    """raised by the shim's tf.placeholder: in the reference's build_model / build_model_ae the first placeholder
    (`self.epoch`) is created right AFTER the losses and the optimizer op -- everything that follows is TensorBoard
    summaries (out of scope), so the build is stopped there."""


def install_training(record):
    """Shim surface for the loss / optimizer wiring of the reference's trainers (trainer.py:160-184,372-396;
    trainer3.py:39-63,255-279): tf.train.AdamOptimizer / GradientDescentOptimizer record their constructor arguments and
    `minimize(loss, global_step, var_list)` records what is minimised over which variables; tf.placeholder stops the build
    (see BuildDone).  `record` is a dict that receives 'optimizer', 'minimize'."""
    tf = sys.modules["tensorflow"]

    class _Opt(object):
        kind = None

        def __init__(self, *args, **kwargs):
            record["optimizer"] = {"kind": self.kind, "args": args, "kwargs": kwargs}

        def minimize(self, loss, global_step=None, var_list=None):
            record["minimize"] = {"loss": loss, "global_step": global_step, "var_list": list(var_list), "optimizer_id": id(self)}
            record.setdefault("minimize_all", []).append(record["minimize"])     # arch=dg minimises twice (d_optim, g_optim)
            return "optim-op"

    train = _mod("tensorflow.train")
    train.AdamOptimizer = type("AdamOptimizer", (_Opt,), {"kind": "adam"})
    train.GradientDescentOptimizer = type("GradientDescentOptimizer", (_Opt,), {"kind": "gd"})
    tf.train = train
    sys.modules["tensorflow.train"] = train

    def placeholder(*a, **k):
        raise BuildDone()

    class Variable(object):
        """tf.Variable used by Trainer.__init__ for `step` and `g_lr` (trainer.py:64,72,76): a named value holder that
        takes part in arithmetic as a tensor of TF's inferred dtype (python int -> int32, python float -> float32)."""

        def __init__(self, initial_value, name=None, trainable=True):
            self.name, self.trainable = name, trainable
            self.value = torch.tensor(initial_value, dtype=torch.int32 if isinstance(initial_value, int) else torch.float32)

        def __mul__(self, other):
            return self.value * other

        __rmul__ = __mul__

    def assign(ref, value, name=None):
        record.setdefault("assign", []).append({"ref": ref, "value": value, "name": name})
        return "assign-op:%s" % name

    tf.Variable, tf.assign = Variable, assign
    tf.cast = lambda x, dtype: (x.value if isinstance(x, Variable) else _t(x)).to(dtype)
    tf.cos = lambda x: torch.cos(_t(x))

    tf.placeholder = placeholder
    tf.square = lambda x: _t(x) ** 2
    tf.sqrt = lambda x: torch.sqrt(_t(x))
    tf.squared_difference = lambda a, b: (_t(a) - _t(b)) ** 2
    tf.reduce_sum = lambda x, axis=None: torch.sum(x) if axis is None else torch.sum(x, dim=axis)

    # tf.distributions.Bernoulli / kl_divergence (use_sparse, trainer.py:389-394): published closed form
    #   KL(Bern(p) || Bern(q)) = p log(p/q) + (1-p) log((1-p)/(1-q))      -- a restatement, like the layer primitives
    ds = _mod("tensorflow.distributions")

    class Bernoulli(object):
        def __init__(self, probs):
            self.probs = torch.as_tensor(probs, dtype=torch.float32)

    def kl_divergence(a, b):
        p, q = a.probs, b.probs
        return p * (p.log() - q.log()) + (1 - p) * ((1 - p).log() - (1 - q).log())

    ds.Bernoulli, ds.kl_divergence = Bernoulli, kl_divergence
    tf.distributions = ds
    sys.modules["tensorflow.distributions"] = ds
    return tf


def import_reference_trainers(reference_root="/root/reference"):
    """Import the reference's trainer.py and trainer3.py verbatim.  `util` (matplotlib / PIL plotting helpers, not
    installable here and never used by build_model) is replaced by an empty module for the import; the visualisation /
    debug helpers the build functions call on the way (`denorm_img`, `denorm_img3`, `show_all_variables`) are replaced by
    no-ops in the imported modules' namespaces -- nothing that enters a loss."""
    import importlib.util

    model = import_reference_model(reference_root)
    ops = import_reference_ops(reference_root)
    saved = {k: sys.modules.get(k) for k in ("ops", "model", "util", "trainer")}
    sys.modules["ops"], sys.modules["model"], sys.modules["util"] = ops, model, types.ModuleType("util")
    mods = {}
    try:
        for fname, key in (("trainer.py", "trainer"), ("trainer3.py", "trainer3")):
            spec = importlib.util.spec_from_file_location("_dfl_reference_" + key, reference_root + "/" + fname)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            mods[key] = mod
            if key == "trainer":
                sys.modules["trainer"] = mod
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    for mod in mods.values():
        mod.denorm_img = lambda *a, **k: None
        mod.denorm_img3 = lambda *a, **k: None
        mod.show_all_variables = lambda: None
    return mods["trainer"], mods["trainer3"], ops


def import_reference_data(reference_root="/root/reference"):
    """Import the reference's data.py verbatim (BatchManager, preprocess).  Its graph-side members are inert stand-ins:
    tf.FIFOQueue / tf.placeholder only have to exist for BatchManager.__init__ (data.py:73-77); matplotlib (imported at
    data.py:12 for the smoke tests at the bottom of the file, not installable here) is an empty module."""
    import importlib.util

    tf = install()

    class FIFOQueue(object):
        def __init__(self, capacity, dtypes, shapes):
            self.capacity, self.dtypes, self.shapes = capacity, dtypes, shapes

        def enqueue(self, vals):
            return "enqueue-op"

        def size(self):
            return 0

    tf.FIFOQueue = FIFOQueue
    keep = getattr(tf, "placeholder", None)
    tf.placeholder = lambda dtype=None, shape=None, name=None: ("placeholder", shape)
    saved = {k: sys.modules.get(k) for k in ("ops", "matplotlib", "matplotlib.pyplot")}
    mpl, plt = types.ModuleType("matplotlib"), types.ModuleType("matplotlib.pyplot")
    mpl.pyplot = plt
    sys.modules["matplotlib"], sys.modules["matplotlib.pyplot"] = mpl, plt
    sys.modules["ops"] = import_reference_ops(reference_root)
    try:
        spec = importlib.util.spec_from_file_location("_dfl_reference_data", reference_root + "/data.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
    mod._restore_placeholder = keep
    return mod


def import_reference_model(reference_root="/root/reference"):
    """Import the reference's model.py verbatim (read-only); it does `from ops import *`, so the reference's ops.py is
    registered under that module name first.  Call install_structural() before using the builders."""
    import importlib.util

    ops = import_reference_ops(reference_root)
    sys.modules["ops"] = ops
    try:
        spec = importlib.util.spec_from_file_location("_dfl_reference_model", reference_root + "/model.py")
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
    finally:
        sys.modules.pop("ops", None)
    return mod


def import_reference_ops(reference_root="/root/reference"):
    """Import the reference's ops.py verbatim (read-only) under the shim."""
    import importlib.util

    install()
    spec = importlib.util.spec_from_file_location("_dfl_reference_ops", reference_root + "/ops.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

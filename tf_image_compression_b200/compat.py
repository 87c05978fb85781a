"""Run the reference's UNMODIFIED entry points (encode.py, decode.py) on the B200 path.

north_star: "the encode.py / decode.py / main.py entry points and the per-model config.json / params layout stay
unchanged".  Those scripts import `tensorflow`, `range_coder`, `skimage.io` and `model_N.model`; none of the first three
exists next to this package and the model modules build TF graphs.  `install(codec)` puts stand-ins for exactly the
surface the two scripts touch into `sys.modules`:

    tensorflow   placeholder / tf.data.Dataset.from_tensor_slices(...).batch().prefetch().make_initializable_iterator()
                 / Session.run(op, feed_dict) / errors.OutOfRangeError (the scripts' end-of-batches control flow,
                 encode.py:160-165, decode.py:215-220) / train.Saver().restore (utils.restore_params,
                 utils/utils.py:84-93 -> checkpoint.restore_params on the TF-V2 bundle) / ConfigProto, set_random_seed.
                 A fetched op is evaluated eagerly: `sess.run(encoder_output_op)` pulls the next <= 64 patches from the
                 iterator and hands them to ModelModule.encoder (the C ABI), exactly one sess.run per batch like the
                 reference.
    range_coder  tf_image_compression_b200.range_coder
    skimage.io   imread / imsave over PIL (PNG)
    model_N.model  ModelModule over the codec (model.encoder(input, patch_size, quan_scale), model.decoder(input, q))

and `load_entry(path)` imports a script of a reference checkout (its own utils/ and data_loader/ packages are imported
unmodified from that checkout).  The scripts' relative paths (model_N/config.json, data_info/distribution_info_N.npy,
model_N/params_for_test/params) are resolved against the current directory, as when the reference runs them.

    codec = T.Codec("model_0", ...)
    with compat.installed(codec, reference_root="/path/to/tf_image_compression"):
        enc = compat.load_entry("encode.py")
        enc.compress(compat.Session(), compat.model_module(0), args)

Anything with the Codec's public methods works as `codec` (tests/oracle_codec.py drives the same scripts with the CPU
oracle to pin the host logic where no GPU is present)."""
from __future__ import annotations

import contextlib
import importlib.util
import sys
import types

import numpy as np

from . import checkpoint, range_coder
from .model_api import ModelModule

_state = {"codec": None, "modules": {}}


class OutOfRangeError(Exception):
    """tf.errors.OutOfRangeError: the iterator is exhausted (control flow in encode.py:164, decode.py:219)."""


class _Placeholder:
    def __init__(self, dtype, shape=None, name=None):
        self.dtype, self.shape, self.name = dtype, shape, name


class _Iterator:
    def __init__(self, source, batch_size):
        self.source, self.batch_size = source, batch_size
        self.data, self.pos = None, 0
        self.initializer = _InitOp(self)

    def get_next(self):
        return _BatchNode(self)

    def next_batch(self):
        if self.data is None or self.pos >= len(self.data):
            raise OutOfRangeError("End of sequence")
        out = self.data[self.pos:self.pos + self.batch_size]
        self.pos += self.batch_size
        return out


class _InitOp:
    def __init__(self, iterator):
        self.iterator = iterator


class _BatchNode:
    def __init__(self, iterator):
        self.iterator = iterator


class _ModelOp:
    """A graph node `model.encoder(batch, ...)` / `model.decoder(batch, ...)`; evaluated by Session.run."""

    def __init__(self, source, fn):
        self.source, self.fn = source, fn


class _Dataset:
    def __init__(self, source):
        self.source, self.batch_size = source, 1

    @staticmethod
    def from_tensor_slices(tensors):
        return _Dataset(tensors)

    def batch(self, batch_size):
        self.batch_size = int(batch_size)
        return self

    def prefetch(self, n):
        return self

    def make_initializable_iterator(self):
        return _Iterator(self.source, self.batch_size)


class Session:
    def __init__(self, config=None, graph=None):
        self.config = config

    def run(self, fetches, feed_dict=None):
        if isinstance(fetches, (list, tuple)):
            return [self.run(f, feed_dict) for f in fetches]
        if isinstance(fetches, _InitOp):
            it = fetches.iterator
            value = (feed_dict or {})[it.source]
            dtype = np.float32 if it.source.dtype in ("float32", np.float32) else None
            it.data = np.asarray(value, dtype=dtype)  # TF converts the fed value to the placeholder's dtype
            it.pos = 0
            return None
        if isinstance(fetches, _ModelOp):
            return fetches.fn(fetches.source.iterator.next_batch())
        raise TypeError(f"compat.Session.run cannot evaluate {type(fetches).__name__}")

    def close(self):
        pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        pass


class _Saver:
    def restore(self, sess, save_path):
        """tf.train.Saver().restore(sess, 'model_N/params_for_test/params'): the TF-V2 bundle -> the codec's graphs."""
        checkpoint.restore_params(_state["codec"], params_file=str(save_path))


class _ConfigProto:
    def __init__(self, **kw):
        self.gpu_options = types.SimpleNamespace(allow_growth=False)
        self.__dict__.update(kw)


class _GraphModel:
    """What `from model_N import model` yields: graph-building encoder / decoder over the lazily evaluated batch."""

    def __init__(self, module: ModelModule):
        self._m = module

    def encoder(self, input, patch_size, quan_scale, bottleneck_channel=None):
        return _ModelOp(input, lambda batch: self._m.encoder(batch, patch_size, quan_scale, bottleneck_channel))

    def decoder(self, input, quan_scale):
        return _ModelOp(input, lambda batch: self._m.decoder(batch, quan_scale))


def model_module(model_num=0, variant=None):
    """The object the scripts get from `from model_N import model` (encode.py:225-232)."""
    codec = _state["codec"]
    if codec is None:
        raise RuntimeError("compat.install(codec) first")
    return _GraphModel(ModelModule(variant or getattr(codec, "variant", f"model_{model_num}"), codec))


def _imread(path):
    from PIL import Image
    return np.asarray(Image.open(path).convert("RGB"))


def _imsave(path, arr):
    from PIL import Image
    Image.fromarray(np.asarray(arr)).save(path)


def _build_modules():
    tf = types.ModuleType("tensorflow")
    tf.float32 = "float32"
    tf.placeholder = _Placeholder
    tf.Session = Session
    tf.ConfigProto = _ConfigProto
    tf.set_random_seed = lambda seed: None
    tf.errors = types.SimpleNamespace(OutOfRangeError=OutOfRangeError)
    tf.data = types.SimpleNamespace(Dataset=_Dataset)
    tf.train = types.SimpleNamespace(Saver=_Saver)
    tf.__path__ = []  # a package: utils/utils.py does `from tensorflow.python.client import timeline`
    py = types.ModuleType("tensorflow.python")
    py.__path__ = []
    client = types.ModuleType("tensorflow.python.client")
    client.__path__ = []
    timeline = types.ModuleType("tensorflow.python.client.timeline")
    client.timeline = timeline
    py.client = client
    tf.python = py
    sk = types.ModuleType("skimage")
    sk.__path__ = []
    io = types.ModuleType("skimage.io")
    io.imread, io.imsave = _imread, _imsave
    sk.io = io
    return {"tensorflow": tf, "tensorflow.python": py, "tensorflow.python.client": client,
            "tensorflow.python.client.timeline": timeline, "skimage": sk, "skimage.io": io, "range_coder": range_coder}


def install(codec, reference_root=None):
    """Register the stand-ins (and the reference checkout on sys.path) for the unmodified scripts."""
    _state["codec"] = codec
    _state["modules"] = _build_modules()
    _state["saved"] = {k: sys.modules.get(k) for k in _state["modules"]}
    sys.modules.update(_state["modules"])
    _state["root"] = str(reference_root) if reference_root else None
    if _state["root"] and _state["root"] not in sys.path:
        sys.path.insert(0, _state["root"])


def uninstall():
    for k, v in _state.get("saved", {}).items():
        if v is None:
            sys.modules.pop(k, None)
        else:
            sys.modules[k] = v
    for k in ("utils", "utils.utils", "data_loader", "data_loader.data_loader"):
        sys.modules.pop(k, None)
    root = _state.get("root")
    if root and root in sys.path:
        sys.path.remove(root)
    _state.update(codec=None, modules={}, saved={}, root=None)


@contextlib.contextmanager
def installed(codec, reference_root=None):
    install(codec, reference_root)
    try:
        yield
    finally:
        uninstall()


def load_entry(path, name=None):
    """Import an unmodified reference script (encode.py, decode.py) as a module without running its __main__ block."""
    name = name or "ref_entry_" + str(path).replace("/", "_").replace(".", "_")
    spec = importlib.util.spec_from_file_location(name, str(path))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod

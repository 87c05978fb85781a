"""B200-native hot path of the tf_image_compression learned codec (sm_100a CUDA behind a C ABI).

Product code.  Never imports oracle/ (test infrastructure) and has no CPU fallback."""
from . import variants  # noqa: F401
from .codec import Codec, TicError, inverse_sigmoid_lut, reference_init  # noqa: F401
from .model_api import ModelModule, load_config, load_normalization  # noqa: F401
from . import utils  # noqa: F401
from . import rmbe  # noqa: F401
from . import parallel, range_coder, entry, checkpoint  # noqa: F401

"""Drop-in import path: `import poisson_CNN` resolves to the B200-native implementation of the
batched-inference hot path (poisson_cnn_b200).  Only the hot-path surface exists:
poisson_CNN.models.{Homogeneous_Poisson_NN_Legacy, Dirichlet_BC_NN_Legacy_2, Poisson_CNN_Legacy},
poisson_CNN.losses.linear_operator_loss, poisson_CNN.convert_tf_object_names."""
import sys

import poisson_cnn_b200 as _impl
from poisson_cnn_b200 import convert_tf_object_names, load_experiment  # noqa: F401
from poisson_cnn_b200 import models, losses, solvers  # noqa: F401

sys.modules[__name__ + ".models"] = models
sys.modules[__name__ + ".losses"] = losses
sys.modules[__name__ + ".solvers"] = solvers

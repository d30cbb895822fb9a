"""Hot-path models with the reference's names (poisson_CNN/models/__init__.py:4-5,9)."""
from .Homogeneous_Poisson_NN_Legacy import Homogeneous_Poisson_NN_Legacy
from .Dirichlet_BC_NN_Legacy import Dirichlet_BC_NN_Legacy_2
from .Poisson_CNN_Legacy import Poisson_CNN_Legacy

# The reference's newer `Homogeneous_Poisson_NN` class cannot be constructed as shipped (it references
# undefined names in __init__); every shipped config uses the Legacy class, so the name aliases it.
Homogeneous_Poisson_NN = Homogeneous_Poisson_NN_Legacy

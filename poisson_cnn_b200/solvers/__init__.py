from .dst import dst_poisson_solve  # noqa: F401
from .pressure_projection import pressure_poisson_solve, neumann_laplacian_apply, hpnn_initial_guess  # noqa: F401

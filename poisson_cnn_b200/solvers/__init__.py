from .dst import dst_poisson_solve  # noqa: F401

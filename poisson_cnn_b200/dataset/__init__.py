"""GPU problem generator: the step before the hot path.

Mirrors the call surface of the reference's poisson_CNN.dataset.generators.numerical
(generate_random_RHS, generate_random_boundaries, numerical_dataset) and
poisson_CNN.dataset.utils.{image_resize, set_max_magnitude_in_batch}; everything runs on the device
through libpcnn kernels (no TensorFlow, no CPU fallback)."""
from .numerical import (generate_random_RHS, generate_random_boundaries, numerical_dataset,  # noqa: F401
                        image_resize, set_max_magnitude_in_batch)

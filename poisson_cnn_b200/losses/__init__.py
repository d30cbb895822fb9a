from .physics_informed_loss import linear_operator_loss  # noqa: F401

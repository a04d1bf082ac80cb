"""Mixed-precision training glue of the reference's `optimization` package that sits on the training hot path:
the loss / gradient scaler (optimization/grad_scaler.py)."""
from .grad_scaler import (AdvancedGradScaler, NumericalError, ScalerConfig, ScalingStrategy, check_gradients_finite,
                          clip_gradients_by_norm, create_scaler, scale_gradients)

GradScaler = AdvancedGradScaler

__all__ = ["AdvancedGradScaler", "GradScaler", "NumericalError", "ScalerConfig", "ScalingStrategy", "check_gradients_finite",
           "clip_gradients_by_norm", "create_scaler", "scale_gradients"]

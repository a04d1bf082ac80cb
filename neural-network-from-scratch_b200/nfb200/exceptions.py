"""The reference's exception vocabulary (exceptions.py:7-424): same class names and hierarchy, the same keyword
arguments feeding `context`, `[ERROR_CODE] message` formatting, `to_dict()`, and the `handle_exception` decorator that
turns stray `ValueError` / `TypeError` / `MemoryError` into `ShapeError` / `DTypeError` / `NumericalError` / `ResourceError`
(error conventions).  Built from one table instead of eleven hand-written classes; every leaf class carries a short list
of suggestions (own wording; upstream's tests look for the API names in them: `reshape()`, `tensor.to(dtype)`, ...).

The package's own error classes (`optim.OptimizerError`, `nn.LayerError`, `models.ModelError`, the scaler's
`NumericalError`) derive from these AND from the builtin they always derived from, so `except ValueError` call sites and the
reference's `except OptimizerError` call sites both keep working."""
from __future__ import annotations

import functools
import traceback
from typing import List, Optional


class NeuralArchError(Exception):
    """Base of the tree: message + machine-readable code + free-form context (exceptions.py:7-68)."""

    def __init__(self, message: str, error_code: Optional[str] = None, context: Optional[dict] = None,
                 suggestions: Optional[List[str]] = None, original_exception: Optional[BaseException] = None) -> None:
        super().__init__(message)
        self.message = message
        self.error_code = error_code or type(self).__name__.upper()
        self.context = dict(context or {})
        self.suggestions = list(suggestions or [])
        self.original_exception = original_exception
        self.stack_trace = traceback.format_stack()

    def __str__(self) -> str:
        parts = [f"[{self.error_code}] {self.message}"]
        if self.context:
            parts.append(f"Context: {self.context}")
        if self.suggestions:
            parts.append("Suggestions:\n" + "\n".join(f"  - {s}" for s in self.suggestions))
        if self.original_exception:
            parts.append(f"Caused by: {self.original_exception}")
        return "\n".join(parts)

    def to_dict(self) -> dict:
        return {"error_type": type(self).__name__, "error_code": self.error_code, "message": self.message, "context": self.context,
                "suggestions": self.suggestions, "original_exception": str(self.original_exception) if self.original_exception else None}


_ADVICE = {
    "ShapeError": ("Check tensor dimensions against what the operation expects", "reshape() or transpose the operand into the expected layout",
                   "Remember the broadcasting rule: trailing axes must be equal or 1"),
    "DTypeError": ("Convert explicitly with tensor.to(dtype) / astype() before the call", "Mixed operands follow NumPy type promotion; state the intended dtype"),
    "DeviceError": ("Bring the operands together with tensor.to(device)", "get_default_device() / set_default_device() decide where new tensors are created",
                    "A cuda tensor needs the b200 backend: there is no silent CPU fallback"),
    "GradientError": ("Create the tensor with requires_grad=True", "Call backward() on a scalar, or pass the upstream gradient explicitly"),
    "NumericalError": ("Look for NaN / inf in the inputs", "Lower the learning rate or clip the gradients"),
    "LayerError": ("Compare the layer's constructor arguments with the input shape",),
    "ParameterError": ("Check the parameter's shape against the layer that owns it",),
    "OptimizerError": ("Check the learning rate and the other hyper-parameters", "Make sure backward() ran and zero_grad() is called once per step"),
    "ConfigurationError": ("Validate the value against the type the option expects",),
    "ResourceError": ("Reduce the batch size or free cached device memory",),
    "DataError": ("Check the format and dtype of the input data",),
}


def _family(name: str, parent: type, code: Optional[str] = None, keys: tuple = ()):
    """A subclass whose keyword arguments (`keys`) land in `context` when given -- the shape of every leaf class upstream."""
    if not keys:
        return type(name, (parent,), {"__doc__": f"{name} (exceptions.py)", "__module__": __name__})

    def __init__(self, message: str, *args, **kwargs):
        given = dict(zip(keys, args))
        given.update({k: v for k, v in kwargs.items() if k in keys})
        unknown = [k for k in kwargs if k not in keys]
        if unknown:
            raise TypeError(f"{name}() got unexpected keyword arguments {unknown}")
        NeuralArchError.__init__(self, message, error_code=code, context={k: v for k, v in given.items() if v is not None},
                                 suggestions=list(_ADVICE.get(name, ())))

    return type(name, (parent,), {"__init__": __init__, "__doc__": f"{name}: context keys {keys} (exceptions.py)", "__module__": __name__})


TensorError = _family("TensorError", NeuralArchError)
ShapeError = _family("ShapeError", TensorError, "SHAPE_MISMATCH", ("expected_shape", "actual_shape", "operation"))
DTypeError = _family("DTypeError", TensorError, "DTYPE_MISMATCH", ("expected_dtype", "actual_dtype"))
DeviceError = _family("DeviceError", TensorError, "DEVICE_MISMATCH", ("expected_device", "actual_device"))
GradientError = _family("GradientError", TensorError, "GRADIENT_ERROR", ("tensor_name", "operation"))
NumericalError = _family("NumericalError", TensorError, "NUMERICAL_ERROR", ("operation", "invalid_values"))
ModelError = _family("ModelError", NeuralArchError)
LayerError = _family("LayerError", ModelError, "LAYER_ERROR", ("layer_name", "layer_type"))
ParameterError = _family("ParameterError", ModelError, "PARAMETER_ERROR", ("parameter_name", "expected_shape", "actual_shape"))
OptimizationError = _family("OptimizationError", NeuralArchError)
OptimizerError = _family("OptimizerError", OptimizationError, "OPTIMIZER_ERROR", ("optimizer_type", "learning_rate"))
ConfigurationError = _family("ConfigurationError", NeuralArchError, "CONFIG_ERROR", ("config_key", "config_value"))
ResourceError = _family("ResourceError", NeuralArchError, "RESOURCE_ERROR", ("resource_type", "required_amount", "available_amount"))
DataError = _family("DataError", NeuralArchError, "DATA_ERROR", ("data_type", "expected_format", "actual_format"))


def handle_exception(func):
    """exceptions.py:376-424: builtin errors escaping `func` come out as members of the tree; tree members pass through."""

    @functools.wraps(func)
    def wrapper(*args, **kwargs):
        try:
            return func(*args, **kwargs)
        except NeuralArchError:
            raise
        except ValueError as e:
            text = str(e).lower()
            if "shape" in text:
                raise ShapeError(f"Shape error in {func.__name__}: {e}", operation=func.__name__) from e
            if "dtype" in text or "type" in text:
                raise DTypeError(f"Data type error in {func.__name__}: {e}") from e
            raise NumericalError(f"Numerical error in {func.__name__}: {e}", operation=func.__name__) from e
        except TypeError as e:
            raise DTypeError(f"Type error in {func.__name__}: {e}") from e
        except MemoryError as e:
            raise ResourceError(f"Memory error in {func.__name__}: {e}", resource_type="memory") from e
        except Exception as e:
            raise NeuralArchError(f"Unexpected error in {func.__name__}: {e}", context={"function": func.__name__}, original_exception=e) from e

    return wrapper

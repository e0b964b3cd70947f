"""Exception types of the detection path.  Interface contract of marEx/exceptions.py:11-120, 180-215,
338-361: the class names, the constructor keywords (``details``, ``suggestions``, ``error_code``,
``context``) and the rendered message layout, which the reference's tests match with regular
expressions.  The rendering itself is a table of sections, written for this package."""
from typing import Any, Callable, Dict, Iterable, List, Optional, Tuple

# (attribute, renderer) in the order the sections appear in the message text
_SECTIONS: Tuple[Tuple[str, Callable[[Any], str]], ...] = (
    ("details", lambda v: "Details: " + str(v)),
    ("context", lambda v: "Context: " + ", ".join(f"{key}={val}" for key, val in v.items())),
    ("suggestions", lambda v: "Suggestions:\n" + "\n".join("  - " + str(item) for item in v)),
    ("error_code", lambda v: "Error Code: " + str(v)),
)


def _render(head: str, fields: Dict[str, Any]) -> str:
    lines: List[str] = [head]
    lines.extend(draw(fields[name]) for name, draw in _SECTIONS if fields.get(name))
    return "\n".join(lines)


class MarExError(Exception):
    """Root of the hierarchy: a headline plus optional details, context, suggestions and a code."""

    default_code: Optional[str] = None

    def __init__(self, message: str, details: Optional[str] = None, suggestions: Optional[Iterable[str]] = None,
                 error_code: Optional[str] = None, context: Optional[Dict[str, Any]] = None):  # fmt: skip
        self.message = message
        self.details = details
        self.suggestions: List[str] = list(suggestions) if suggestions else []
        self.error_code = error_code if error_code is not None else self.default_code
        self.context: Dict[str, Any] = dict(context) if context else {}
        super().__init__(self._format_error_message())

    def _format_error_message(self) -> str:
        return _render(self.message, vars(self))

    def add_suggestion(self, suggestion: str) -> None:
        self.suggestions += [suggestion]

    def add_context(self, key: str, value: Any) -> None:
        self.context[key] = value


class DataValidationError(MarExError):
    """The input data cannot be processed (missing dims / coords, no finite data, too few years ...)."""

    default_code = "DATA_VALIDATION"


class ConfigurationError(MarExError):
    """A parameter, or a combination of parameters, is not valid."""

    default_code = "CONFIGURATION_ERROR"


class ProcessingError(MarExError):
    """libmarex_b200 reported a failure (a negative MAREX_ERR_* return code) or is not built."""

    default_code = "PROCESSING_ERROR"


def create_data_validation_error(message: str, data_info: Optional[Dict[str, Any]] = None, **kwargs) -> DataValidationError:
    """Factory used by the validators: ``data_info`` is merged into the error's context."""
    merged = dict(kwargs.pop("context", None) or {})
    merged.update(data_info or {})
    return DataValidationError(message, context=merged, **kwargs)

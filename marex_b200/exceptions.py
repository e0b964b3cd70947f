"""Exception types of the detection path, mirroring marEx/exceptions.py:11-120, 180-215, 338-361:
same class names, constructor arguments and message layout (the reference's tests match on the
message text)."""
from typing import Any, Dict, List, Optional


class MarExError(Exception):
    """Base class (marEx/exceptions.py:11-82)."""

    def __init__(
        self,
        message: str,
        details: Optional[str] = None,
        suggestions: Optional[List[str]] = None,
        error_code: Optional[str] = None,
        context: Optional[Dict[str, Any]] = None,
    ):
        self.message = message
        self.details = details
        self.suggestions = suggestions or []
        self.error_code = error_code
        self.context = context or {}
        super().__init__(self._format_error_message())

    def _format_error_message(self) -> str:
        parts = [self.message]
        if self.details:
            parts.append(f"Details: {self.details}")
        if self.context:
            parts.append("Context: " + ", ".join(f"{k}={v}" for k, v in self.context.items()))
        if self.suggestions:
            parts.append("Suggestions:\n" + "\n".join(f"  - {s}" for s in self.suggestions))
        if self.error_code:
            parts.append(f"Error Code: {self.error_code}")
        return "\n".join(parts)

    def add_suggestion(self, suggestion: str) -> None:
        self.suggestions.append(suggestion)

    def add_context(self, key: str, value: Any) -> None:
        self.context[key] = value


class DataValidationError(MarExError):
    """Input data problems (marEx/exceptions.py:84-120)."""

    def __init__(self, message, details=None, suggestions=None, error_code="DATA_VALIDATION", context=None):
        super().__init__(message, details, suggestions, error_code, context)


class ConfigurationError(MarExError):
    """Invalid parameter combinations (marEx/exceptions.py:180-215)."""

    def __init__(self, message, details=None, suggestions=None, error_code="CONFIGURATION", context=None):
        super().__init__(message, details, suggestions, error_code, context)


class ProcessingError(MarExError):
    """Failures inside the CUDA library (negative MAREX_ERR_* return codes)."""

    def __init__(self, message, details=None, suggestions=None, error_code="PROCESSING", context=None):
        super().__init__(message, details, suggestions, error_code, context)


def create_data_validation_error(message: str, data_info: Optional[Dict[str, Any]] = None, **kwargs) -> DataValidationError:
    """marEx/exceptions.py:338-361."""
    context = kwargs.get("context", {})
    if data_info:
        context.update(data_info)
    kwargs["context"] = context
    return DataValidationError(message, **kwargs)

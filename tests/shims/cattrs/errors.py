"""cattrs.errors stand-in (TEST INFRASTRUCTURE ONLY)."""
from cattrs import (BaseValidationError, ClassValidationError, ForbiddenExtraKeysError,  # noqa: F401
                    IterableValidationError, StructureHandlerNotFoundError)

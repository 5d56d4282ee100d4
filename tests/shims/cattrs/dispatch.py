"""cattrs.dispatch stand-in (TEST INFRASTRUCTURE ONLY)."""
from typing import Any, Callable

UnstructureHook = Callable[[Any], Any]
StructureHook = Callable[[Any, Any], Any]

"""Controller registry (reference: src/controllers/__init__.py:1-7)."""
from .basic_controller import BasicMAC
from .entity_controller import EntityMAC

REGISTRY = {"basic_mac": BasicMAC, "entity_mac": EntityMAC}

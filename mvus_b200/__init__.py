"""mvus-b200: B200-native bundle adjustment for CenekAlbl/mvus (drop-in for Scene.BA)."""
from .scene import Camera, Scene  # noqa: F401

__all__ = ['Scene', 'Camera']

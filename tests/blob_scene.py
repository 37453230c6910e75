"""Rebuilds a Scene (host tables) from a packed blob, so that tests on the GPU box -- where
/root/reference and its XML/mesh assets do not exist -- can drive the numpy oracle from the
committed golden fixtures."""
from tactilesimulation_b200.layout import scene_from_blob  # noqa: F401

"""``KnowledgeGraph`` under the reference's import path (``torchkge.data_structures``); the class
itself lives in ``torchkge_b200.data`` and is the index-tensor view the hot path reads."""
from .data import FilterIndex, KnowledgeGraph  # noqa: F401

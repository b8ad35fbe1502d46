"""Name-compatible home of the SSL trainers (the reference keeps them in pretrainer.py)."""
from .trainer import DifHeadTrainer, EdgeLabels, GeneratedEdgeTrainer, SupEdgeTrainer  # noqa: F401

SSL_TRAINERS = {"DisEdge": GeneratedEdgeTrainer, "SupEdge": SupEdgeTrainer, "DifHead": DifHeadTrainer}

"""qsv — B200-native state-vector engine behind the Quantum-Circuits API."""
from . import _cabi, schedule  # noqa: F401  (engine/lowering import torch lazily via their users)

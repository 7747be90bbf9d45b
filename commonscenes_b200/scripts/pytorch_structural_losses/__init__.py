"""Mirror of the reference's `scripts/pytorch_structural_losses` package (its `StructuralLossesBackend` pybind module,
src/structural_loss.cpp:127-133, is replaced by the C ABI of libcsb200.so): `nn_distance`, `match_cost`."""
from .match_cost import match_cost          # noqa: F401
from .nn_distance import nn_distance        # noqa: F401

"""Trainer3 (reference trainer3.py:13-368): the 3D overrides.  In the reference only the graph construction differs
(GeneratorBE3, curl = jacobian3(G_s)[1], jacobian3 loss: trainer3.py:14-63); here the engine and the fused stencil
kernel are dimension-generic, so Trainer3 only pins `is_3d` semantics.  The reference's `build_test_model` bug
(2D `curl` applied to the 3-channel potential, trainer3.py:188) is NOT reproduced: the 3D curl is used everywhere."""
from .trainer import Trainer


class Trainer3(Trainer):
    def __init__(self, config, batch_manager):
        assert config.is_3d, "Trainer3 is the 3D trainer (main.py:20-23)"
        super(Trainer3, self).__init__(config, batch_manager)

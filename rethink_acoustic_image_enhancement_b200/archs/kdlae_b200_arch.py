"""BasicSR registry shim (Train/basicsr/models/archs/__init__.py:9-46).

`define_network(opt)` scans every `*_arch.py` in basicsr/models/archs/ and instantiates
`getattr(module, opt['type'])(**opt)`.  Copy or symlink this file into that folder (see
INTEGRATION.md) and the yaml `network_g.type: RestormerSuperResolutionParam2 | KDLAE_teacher |
KDLAE_student` resolves to the B200-native modules with unchanged kwargs.
"""
from rethink_acoustic_image_enhancement_b200.kdlae_model import (  # noqa: F401
    KDLAE_teacher,
    KDLAE_student,
    RestormerSuperResolutionParam2,
)

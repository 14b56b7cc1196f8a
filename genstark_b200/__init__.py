"""genstark_b200 -- B200-native STARK proving hot path behind genSTARK's instantiate()/prove()/verify() API."""
from .air import AirModule, ProgramBuilder, StaticRegister, P128, P32  # noqa: F401
from . import airs  # noqa: F401


def instantiate(source, component='default', options=None, logger=None, context=None):
    """index.ts:18-33: AirAssembly text / path of an .aa file / AirModule -> Stark"""
    from .stark import instantiate as _inst
    return _inst(source, component, options, logger, context)

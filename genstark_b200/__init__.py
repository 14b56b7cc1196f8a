"""genstark_b200 -- B200-native STARK proving hot path behind genSTARK's instantiate()/prove()/verify() API."""
from .air import AirModule, ProgramBuilder, StaticRegister, P128, P32  # noqa: F401
from . import airs  # noqa: F401


def instantiate(air, options=None, logger=None):
    from .stark import instantiate as _inst
    return _inst(air, options, logger)

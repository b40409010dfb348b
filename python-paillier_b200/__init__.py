"""python-paillier_b200: a B200-native batched Paillier engine behind the ``phe`` API.

The directory name carries a hyphen (it mirrors the reference repo's name), so import it with
``importlib.import_module("python-paillier_b200")`` or through the root-level alias module
``paillier_b200`` (``import paillier_b200 as phe``).
"""
from .engine import (Engine, EngineError, EngineUnavailable, ModContext, PrivateContext, PublicContext,  # noqa: F401
                     get_engine, ints_to_limbs, limbs_to_ints)

__version__ = "0.1.0"

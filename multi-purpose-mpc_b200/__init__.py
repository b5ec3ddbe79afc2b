"""multi-purpose-mpc_b200 -- B200-native batched closed-loop MPC engine behind the Python API of
matssteinweg/Multi-Purpose-MPC (Map, ReferencePath, BicycleModel, MPC).

The directory name contains hyphens, so import it through the root-level loader: `import mpc_b200`.
"""
from ._lib import Engine, MpcError, MpcConfig, default_config, path_table, load, LIB_PATH, EXPORTS  # noqa: F401
from ._lib import (ST_QP_FALLBACK, ST_DEAD, ST_NO_SEGMENT, ST_END_OF_PATH, ST_INDEX_ERROR, ST_FINISHED)  # noqa: F401

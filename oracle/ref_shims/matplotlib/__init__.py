"""No-op stand-in so the reference's modules import (matplotlib is visualisation only)."""

"""Cell sequence of an anti-aliased line (Zingl's algorithm as used by skimage.draw.line_aa), host
side.  Only Map.add_boundary needs it on the host (setup); the raycast kernels walk the same
sequence on the device (csrc/geometry.cu::line_aa_walk)."""
import math

import numpy as np


def line_aa_cells(r0, c0, r1, c1):
    """Yield the (r, c) cells in emission order for the line (r0, c0) -> (r1, c1)."""
    dc, dr = abs(c0 - c1), abs(r0 - r1)
    err = np.float32(dc - dr)
    step_c = 1 if c0 < c1 else -1
    step_r = 1 if r0 < r1 else -1
    ed = np.float32(1.0) if dc + dr == 0 else np.float32(math.sqrt(dc * dc + dr * dr))
    r, c = r0, c0
    while True:
        yield r, c
        e0, c_start = err, c
        if 2 * e0 >= -dc:
            if c == c1:
                return
            if e0 + np.float32(dr) < ed:
                yield r + step_r, c
            err = np.float32(err - dr)
            c += step_c
        if 2 * e0 <= dr:
            if r == r1:
                return
            if np.float32(dc) - e0 < ed:
                yield r, c_start + step_c
            err = np.float32(err + dc)
            r += step_r

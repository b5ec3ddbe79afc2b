"""Map / Obstacle -- host mirror of the reference's src/map.py (same names, arguments, attributes).

`Map.data` keeps the reference's layout (H x W int8, 1 = free, 0 = occupied, row index = world-y
cell, map.py:61-114) because user code indexes it directly.  The CUDA engine holds a bit-packed copy
(engine.set_base_grid) that is refreshed whenever `data` changes through this class; per-scenario
obstacle sets of the batched API are rasterised on the device (mpc_set_obstacles).
"""
import numpy as np


class Obstacle:
    """Circular obstacle in world coordinates (map.py:16-26)."""

    def __init__(self, cx, cy, radius):
        self.cx = cx
        self.cy = cy
        self.radius = radius

    def show(self):  # visualisation is out of scope (matplotlib is not a dependency)
        raise NotImplementedError("plotting is not part of the B200 engine")


def _fill_small_holes(free, area_threshold, connectivity):
    """What skimage.morphology.remove_small_holes(ar, area_threshold, connectivity) does to a 2-D
    boolean image: background components with fewer than area_threshold pixels become foreground."""
    from scipy import ndimage
    free = np.asarray(free, dtype=bool)
    structure = ndimage.generate_binary_structure(free.ndim, min(int(connectivity), free.ndim))
    labels, _ = ndimage.label(~free, structure=structure)
    counts = np.bincount(labels.ravel())
    small = counts < area_threshold
    small[0] = False
    filled = free.copy()
    filled[small[labels]] = True
    return filled


class Map:
    def __init__(self, file_path, origin, resolution, threshold_occupied=100):
        """Occupancy grid loaded from an image (channel 0), binarised at threshold_occupied
        (map.py:45-75).  `file_path` may also be a 2-D array of raw grey levels."""
        self.threshold_occupied = threshold_occupied
        if isinstance(file_path, np.ndarray):
            raw = file_path
        else:
            from PIL import Image
            img = np.array(Image.open(file_path))
            raw = img[:, :, 0] if img.ndim == 3 else img
        self.data = raw
        self.process_map()
        self.height = self.data.shape[0]
        self.width = self.data.shape[1]
        self.resolution = resolution
        self.origin = origin
        self.obstacles = list()
        self.boundaries = list()
        self.version = 0  # bumped on every change of `data`; engines re-upload when it moves

    @classmethod
    def from_grid(cls, data, origin, resolution):
        """Build a Map from an already binarised grid (1 free / 0 occupied)."""
        m = cls.__new__(cls)
        m.threshold_occupied = 1
        m.data = np.ascontiguousarray(data, dtype=np.int8)
        m.height, m.width = m.data.shape
        m.resolution, m.origin = resolution, origin
        m.obstacles, m.boundaries, m.version = [], [], 0
        return m

    def w2m(self, x, y):
        """World -> cell (map.py:77-88): fp64 divide then floor."""
        dx = int(np.floor((x - self.origin[0]) / self.resolution))
        dy = int(np.floor((y - self.origin[1]) / self.resolution))
        return dx, dy

    def m2w(self, dx, dy):
        """Cell -> world coordinates of the cell centre (map.py:90-101)."""
        x = (dx + 0.5) * self.resolution + self.origin[0]
        y = (dy + 0.5) * self.resolution + self.origin[1]
        return x, y

    def process_map(self):
        """Binarise and fill holes smaller than 5 px, 8-connected (map.py:103-114)."""
        free = np.where(self.data >= self.threshold_occupied, 1, 0)
        self.data = _fill_small_holes(free, area_threshold=5, connectivity=8).astype(np.int8)

    def add_obstacles(self, obstacles):
        """Rasterise discs into `data` (map.py:116-137): radius_px = ceil(r / res), window
        [c - r, c + r) in both axes, cells with dx^2 + dy^2 <= r^2 become occupied."""
        self.obstacles.extend(obstacles)
        for ob in obstacles:
            r = int(np.ceil(ob.radius / self.resolution))
            cx, cy = self.w2m(ob.cx, ob.cy)
            yy, xx = np.ogrid[-r:r, -r:r]
            inside = xx ** 2 + yy ** 2 <= r ** 2
            self.data[cy - r:cy + r, cx - r:cx + r][inside] = 0
        self.version += 1

    def add_boundary(self, boundaries):
        """Rasterise anti-aliased line boundaries (map.py:139-155)."""
        from .line_aa import line_aa_cells
        self.boundaries.extend(boundaries)
        for b in boundaries:
            sx = self.w2m(b[0][0], b[0][1])
            gx = self.w2m(b[1][0], b[1][1])
            for x, y in line_aa_cells(sx[0], sx[1], gx[0], gx[1]):
                self.data[y, x] = 0
        self.version += 1

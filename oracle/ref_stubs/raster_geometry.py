"""TEST INFRASTRUCTURE ONLY -- `raster_geometry.cylinder` / `cube` RESTATED (PARITY UNPINNED, see README.md).

Call sites in the reference: ctunet/utilities.py:18, 161-163 (keyword arguments ``shape, height, radius, axis,
position`` and ``shape, side, position``, positions relative).  The published algorithm: ``coord(shape, position,
is_relative=True, use_int=True)`` turns the relative position into the absolute origin ``round((dim - 1) * rel)``
and returns the open grid of integer offsets from it; a cube is ``|offset| <= side / 2`` on every axis, a cylinder
a disk ``sum of squared offsets <= radius^2`` in the two other axes extruded over ``|offset| <= height / 2``."""
import numpy as np


def coord(shape, position=0.5, is_relative=True, use_int=True):
    if not hasattr(position, "__len__"):
        position = (position,) * len(shape)
    if is_relative:
        origin = tuple((dim - 1.0) * rel for dim, rel in zip(shape, position))
    else:
        origin = tuple(position)
    if use_int:
        origin = tuple(int(round(x)) for x in origin)
    return np.ogrid[tuple(slice(-x0, dim - x0) for x0, dim in zip(origin, shape))]


def cube(shape, side, position=0.5):
    if not hasattr(shape, "__len__"):
        shape = (shape,) * 3
    rendered = np.ones(tuple(shape), dtype=bool)
    for x_i in coord(shape, position):
        rendered = rendered * (np.abs(x_i) <= side / 2.0)
    return rendered


def cylinder(shape, height, radius, axis=-1, position=0.5):
    if not hasattr(shape, "__len__"):
        shape = (shape,) * 3
    xx = coord(shape, position)
    axis = axis % 3
    r2 = sum(xx[i].astype(np.float64) ** 2 for i in range(3) if i != axis)
    return (r2 <= radius ** 2) * (np.abs(xx[axis]) <= height / 2.0)

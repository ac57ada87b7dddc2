"""Host-side mirror of the framebuffer tile partition (csrc/wavefront.cuh: LocalPixelToImage).

The frame is cut into 8x8-pixel tiles numbered row-major; rank r of `world` owns tiles r, r+world, ...
(interleaved so the empty background of a scene like Dragon is balanced).  Each rank
renders into a compact tile buffer of `pixels_per_rank * 3` floats — equal on every rank, so ONE
all-gather moves every rank's buffer — and `assemble` scatters the gathered buffers to the row-major
frame.  The CUDA path does the same scatter in k_assemble; this numpy version is what the CPU (gloo)
tests and host-side tools use.
"""
import numpy as np

TILE = 8
TILE_PIXELS = TILE * TILE


def num_tiles(width, height):
    return ((width + TILE - 1) // TILE) * ((height + TILE - 1) // TILE)


def pixels_per_rank(width, height, world):
    return ((num_tiles(width, height) + world - 1) // world) * TILE_PIXELS


def local_pixel_to_image(width, height, world, rank, local_pixel):
    """Vectorised: returns (i, j, valid) for local pixel indices of `rank`."""
    local_pixel = np.asarray(local_pixel, dtype=np.int64)
    tiles_x = (width + TILE - 1) // TILE
    tile = (local_pixel // TILE_PIXELS) * world + rank
    in_tile = local_pixel % TILE_PIXELS
    i = (tile % tiles_x) * TILE + in_tile % TILE
    j = (tile // tiles_x) * TILE + in_tile // TILE
    valid = (tile < num_tiles(width, height)) & (i < width) & (j < height)
    return i, j, valid


def extract(frame, world, rank):
    """Row-major frame [h, w, 3] -> this rank's tile buffer [pixels_per_rank, 3] (padding pixels are 0)."""
    h, w = frame.shape[:2]
    n = pixels_per_rank(w, h, world)
    i, j, valid = local_pixel_to_image(w, h, world, rank, np.arange(n))
    out = np.zeros((n, 3), dtype=frame.dtype)
    out[valid] = frame[j[valid], i[valid]]
    return out


def assemble(gathered, width, height):
    """gathered [world, pixels_per_rank, 3] -> row-major frame [height, width, 3]."""
    world = gathered.shape[0]
    frame = np.zeros((height, width, 3), dtype=gathered.dtype)
    n = gathered.shape[1]
    for rank in range(world):
        i, j, valid = local_pixel_to_image(width, height, world, rank, np.arange(n))
        frame[j[valid], i[valid]] = gathered[rank][valid]
    return frame

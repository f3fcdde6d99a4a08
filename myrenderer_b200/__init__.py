"""myrenderer_b200 -- B200-native geometry generation for platypro/myrenderer's hot path.

The product is the C-ABI library `lib/libmyrenderer_b200.so` (CUDA kernels for sm_100a, declared in
include/myrenderer_b200.h).  This package is the thin host layer over it, mirroring the reference's
Terrain / Polygon / Triangulation / VertexBuffer / VertexLayout / unirand interface.
"""
from ._capi import (LIB_PATH, MrError, MR_POLY_ARENA, MR_POLY_DEGENERATE, MR_POLY_NONFINITE,
                    MR_POLY_NULL_UNWRAP, MR_POLY_OK, MR_POLY_OVERFLOW, MR_POLY_STUCK,
                    MR_POLY_TOO_LARGE, MR_POLY_UNDERFILL, load)
from .host import (Context, GPUVertex, Polygon, PolygonBatch, PolygonObj, Terrain, TerrainMesh,
                   TerrainVertex, Triangulation, Unirand, VertexBuffer, VertexLayout,
                   load_heightmap_png, polygon_offsets_host, unirand_seed)

__all__ = [
    "Context", "Terrain", "TerrainMesh", "Polygon", "PolygonBatch", "PolygonObj", "Triangulation",
    "VertexBuffer", "VertexLayout", "GPUVertex", "TerrainVertex", "Unirand", "unirand_seed",
    "load_heightmap_png", "polygon_offsets_host", "MrError", "load", "LIB_PATH",
]

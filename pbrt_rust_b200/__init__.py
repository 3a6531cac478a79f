"""pbrt_rust_b200 — B200-native (sm_100a) rendering back end behind pbrt_rust's
SamplerRenderer::render.  The product is libpbrtb200.so (hand-written CUDA kernels + C ABI, see
include/pbrtb200.h); this package is the ctypes binding and a data-only mirror of the reference's
scene/camera/sampler constructors.  There is no CPU fallback: the library must be built and a CUDA
device must be present to render."""
from ._ffi import LIB_PATH, MISS, lib
from .api import (HIT_DTYPE, AreaLight, Camera, Context, DevicePtr, Film, Filter, GpuRenderer, Group, HostScene, Light,
                  Material, PbrtError, PlanarMapping2D, Primitive, Sampler, Scene, Shape,
                  SurfaceIntegrator, Texture, Transform, UVMapping2D, film_to_rgb, rgb_to_bytes, write_image)

__all__ = [n for n in dir() if not n.startswith("_")]

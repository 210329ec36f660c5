// Incremental walk over the tiles first + i * stride of a persistent CTA.  A tile index decomposes into (image b, tile row ty, tile
// column tx); spelled as tile / tiles_per_image, rem / tiles_x, rem % tiles_x (and i % stages for the ring) every warp paid three
// integer divisions per tile - on short-range block 1 ~170 of the ~680 instructions a compute warp issues per tile (ncu source
// counters).  Here the decomposition is done ONCE for the first tile and once for the stride; each step is three adds and two carries.
#pragma once

#ifdef __CUDACC__
#define FDT_HD __host__ __device__ __forceinline__
#else
#define FDT_HD inline
#endif

namespace fdt {

struct TileAt { int b, ty, tx; };
struct TileStep { int b, ty, tx; };

FDT_HD TileAt tile_at(int tile, int tpi, int tiles_x) {
  TileAt a;
  a.b = tile / tpi;
  const int r = tile - a.b * tpi;
  a.ty = r / tiles_x; a.tx = r - a.ty * tiles_x;
  return a;
}
FDT_HD TileStep tile_step(int stride, int tpi, int tiles_x) {
  const TileAt a = tile_at(stride, tpi, tiles_x);
  TileStep s; s.b = a.b; s.ty = a.ty; s.tx = a.tx;
  return s;
}
FDT_HD void tile_advance(TileAt& a, const TileStep& s, int tiles_x, int tiles_y) {
  a.tx += s.tx; a.ty += s.ty; a.b += s.b;
  if (a.tx >= tiles_x) { a.tx -= tiles_x; ++a.ty; }
  if (a.ty >= tiles_y) { a.ty -= tiles_y; ++a.b; }
}

}  // namespace fdt

// dfilm.cuh — film accumulation without atomics: every film pixel GATHERS the samples whose filter footprint
// covers it, in a fixed order, then adds the wave's partial sum to its {X,Y,Z,sum_w} record.
// Replaces reference src/core/film/film_tile.rs:84-183 (FilmTile::add_sample_filter: 16x16 filter table,
// per-sample weights normalised to sum 1), src/core/film/film.rs:219-241 (merge_film_tile: Spectrum -> XYZ, +=)
// and film.rs:440-472 (write_image: XYZ -> RGB / sum_w, clamp, scale).
#pragma once
#include "dshade.cuh"

namespace ptd {

PT_HD int film_fidx(float d, float radius, float l) {  // filter-table index along one axis, -1 = outside
  if (!(d <= radius)) return -1;
  int id = (int)floorf(d * l);
  return id < 15 ? id : 15;
}

// One film pixel (fx, fy) of the cropped pixel bounds, for a wave holding `ns` samples of each pixel listed in `slot_of`
// (slot_of[(y - sb.y0) * sbw + (x - sb.x0)] = slot or -1); path id = base + s * npix + slot (base = first path slot of the
// wave's segment of the state).
template <int NC>
PT_HD void film_gather_body(const DScene& S, const WaveState& W, const int* slot_of, uint32_t base, uint32_t npix, uint32_t ns, int fx, int fy, f4* film) {
  const pt_film& F = S.film;
  const float rx = F.filter_radius[0], ry = F.filter_radius[1];
  const float lx = (1.0f / rx) * 15.0f, ly = (1.0f / ry) * 15.0f;
  const int* sb = F.sample_bounds;
  const int* cb = F.cropped_pixel_bounds;
  const int sbw = sb[2] - sb[0];
  // sample pixels q whose samples (q + u, u in [0,1)) can reach this pixel centre: |f + 0.5 - (q + u)| <= r
  int qx0 = (int)ceilf((float)fx + 0.5f - rx - 1.0f), qx1 = (int)floorf((float)fx + 0.5f + rx);
  int qy0 = (int)ceilf((float)fy + 0.5f - ry - 1.0f), qy1 = (int)floorf((float)fy + 0.5f + ry);
  if (qx0 < sb[0]) qx0 = sb[0];
  if (qy0 < sb[1]) qy0 = sb[1];
  if (qx1 > sb[2] - 1) qx1 = sb[2] - 1;
  if (qy1 > sb[3] - 1) qy1 = sb[3] - 1;
  Sp<NC> acc = sp_const<NC>(0.0f);
  float wacc = 0.0f;
  bool any = false;
  for (int qy = qy0; qy <= qy1; qy++)
    for (int qx = qx0; qx <= qx1; qx++) {
      int slot = slot_of[(qy - sb[1]) * sbw + (qx - sb[0])];
      if (slot < 0) continue;
      for (uint32_t s = 0; s < ns; s++) {
        uint32_t i = base + s * npix + (uint32_t)slot;
        f2 pf = W.p_film[i];
        int ix = film_fidx(fabsf((float)fx + 0.5f - pf.x), rx, lx);
        int iy = film_fidx(fabsf((float)fy + 0.5f - pf.y), ry, ly);
        if (ix < 0 || iy < 0) continue;
        // this sample's footprint and weight normalisation (film_tile.rs:96-160)
        int p0x = (int)floorf(pf.x - rx), p0y = (int)floorf(pf.y - ry);
        int p1x = (int)ceilf(pf.x + rx), p1y = (int)ceilf(pf.y + ry);
        if (p0x < cb[0]) p0x = cb[0];
        if (p0y < cb[1]) p0y = cb[1];
        if (p1x > cb[2]) p1x = cb[2];
        if (p1y > cb[3]) p1y = cb[3];
        float sum = 0.0f;
        for (int y = p0y; y < p1y; y++) {
          int jy = film_fidx(fabsf((float)y + 0.5f - pf.y), ry, ly);
          for (int x = p0x; x < p1x; x++) {
            int jx = film_fidx(fabsf((float)x + 0.5f - pf.x), rx, lx);
            float w = (jx < 0 || jy < 0) ? 0.0f : F.filter_table[jy * 16 + jx];
            sum += w;
          }
        }
        if (sum <= 0.0f) continue;
        float w = F.filter_table[iy * 16 + ix] * (1.0f / sum);
        Sp<NC> L = st_load<NC>(W.L, W.cap, i);
        float y = sp_y(S, L);
        if (y > F.max_sample_luminance) L = L * (F.max_sample_luminance / y);
        acc = acc + L * 1.0f * w;
        wacc += w;
        any = true;
      }
    }
  if (!any) return;
  float xyz[3];
  sp_xyz(S, acc, xyz);
  size_t fi = (size_t)(fy - cb[1]) * (size_t)(cb[2] - cb[0]) + (size_t)(fx - cb[0]);
  f4 p = film[fi];
  p.x += xyz[0]; p.y += xyz[1]; p.z += xyz[2]; p.w += wacc;
  film[fi] = p;
}

PT_HD void film_resolve_body(const pt_film& F, f4 p, float* rgb) {  // film.rs:447-472
  float c0 = 3.240479f * p.x - 1.537150f * p.y - 0.498535f * p.z;
  float c1 = -0.969256f * p.x + 1.875991f * p.y + 0.041556f * p.z;
  float c2 = 0.055648f * p.x - 0.204043f * p.y + 1.057311f * p.z;
  if (p.w > 0.0f) {
    float inv = 1.0f / p.w;
    c0 = fmaxf(0.0f, c0 * inv); c1 = fmaxf(0.0f, c1 * inv); c2 = fmaxf(0.0f, c2 * inv);
  }
  rgb[0] = (c0 + 0.0f) * F.scale; rgb[1] = (c1 + 0.0f) * F.scale; rgb[2] = (c2 + 0.0f) * F.scale;
}

}  // namespace ptd

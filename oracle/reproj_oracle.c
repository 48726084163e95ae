/*
 * ORACLE — test infrastructure only.  CPU restatement of the integer part of EvoWorld's
 * reprojection path.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file; the product never does.
 *
 * What it restates
 *   CubemapRenderer.render_face / render_cubemap  evoworld/reprojection/reproject_vggt_open3d_utils.py:617-666
 *   (Open3D 0.18 OffscreenRenderer point pass — un-vendored; parity UNPINNED):
 *     for every (view, face): X_c = W2C [X_w 1];  keep z > z_near;
 *     u = fx x/z + cx, v = fy y/z + cy (fx=fy=res/2, cx=cy=res/2, :624-628);
 *     pixel (floor u, floor v) inside [0,res)^2; nearest z wins, ties -> lowest point index.
 *   The arithmetic (fmaf chain, IEEE division) is fixed so that the CUDA kernel
 *   (evoworld_b200/csrc/reproj.cu:splat_one) reproduces it bit for bit.
 *
 * Build: gcc -O2 -fPIC -shared -std=c11 -ffp-contract=off -mfma -fopenmp (evoworld_b200/build.py)
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static inline uint32_t f2u(float f) { uint32_t u; memcpy(&u, &f, 4); return u; }

/* pts4 [n,4] float {x,y,z,rgbbits}; w2c [V,6,12]; keys [V,6,res,res] uint64 (all-ones = empty) */
void oracle_splat_keys(const float* pts4, int64_t n, const float* w2c, int V, int res, float focal,
                       float z_near, uint64_t* keys) {
  const float c = 0.5f * (float)res;
  const float fres = (float)res;
  const int64_t face_sz = (int64_t)res * res;
#pragma omp parallel for schedule(dynamic, 1)
  for (int vf = 0; vf < V * 6; ++vf) {
    const float* m = w2c + (int64_t)vf * 12;
    uint64_t* zf = keys + (int64_t)vf * face_sz;
    for (int64_t i = 0; i < face_sz; ++i) zf[i] = UINT64_MAX;
    for (int64_t i = 0; i < n; ++i) {
      const float x = pts4[i * 4], y = pts4[i * 4 + 1], z = pts4[i * 4 + 2];
      float zc = fmaf(m[8], x, fmaf(m[9], y, fmaf(m[10], z, m[11])));
      if (!(zc > z_near)) continue;
      float xc = fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
      float yc = fmaf(m[4], x, fmaf(m[5], y, fmaf(m[6], z, m[7])));
      float u = fmaf(focal, xc / zc, c);
      float v = fmaf(focal, yc / zc, c);
      if (!(u >= 0.f && u < fres && v >= 0.f && v < fres)) continue;
      int px = (int)floorf(u), py = (int)floorf(v);
      uint64_t key = ((uint64_t)f2u(zc) << 32) | (uint32_t)i;
      uint64_t* cell = zf + (int64_t)py * res + px;
      if (key < *cell) *cell = key;
    }
  }
}

/* Cube formulation (evoworld_b200/csrc/reproj.cu:cube_splat_one): one camera-space transform per (point, view) with
 * the FRONT camera's w2c [V,12]; face = major axis, face-local coordinates = inv(T_face [Rz180]) applied exactly:
 *   front (x,y,z) | right (-z,y,x) | back (-x,y,-z) | left (z,y,-x) | top (-x,-z,-y) | bottom (-x,z,y)
 * (CUBEMAP_TRANSFORMS, reproject_vggt_open3d_utils.py:29-36; do_flip Rz(180) for top/bottom :619-622,652-655). */
static void splat_keys_cube_core(const float* pts4, int64_t n, const float* w2c_front, int V, int res, float focal,
                                 float z_near, uint64_t* keys, int color_keys) {
  const float c = 0.5f * (float)res;
  const float fres = (float)res;
  const int64_t face_sz = (int64_t)res * res;
#pragma omp parallel for schedule(dynamic, 1)
  for (int v = 0; v < V; ++v) {
    const float* m = w2c_front + (int64_t)v * 12;
    uint64_t* zv = keys + (int64_t)v * 6 * face_sz;
    for (int64_t i = 0; i < 6 * face_sz; ++i) zv[i] = UINT64_MAX;
    for (int64_t i = 0; i < n; ++i) {
      const float x = pts4[i * 4], y = pts4[i * 4 + 1], z = pts4[i * 4 + 2];
      const float xc = fmaf(m[0], x, fmaf(m[1], y, fmaf(m[2], z, m[3])));
      const float yc = fmaf(m[4], x, fmaf(m[5], y, fmaf(m[6], z, m[7])));
      const float zc = fmaf(m[8], x, fmaf(m[9], y, fmaf(m[10], z, m[11])));
      const float ax = fabsf(xc), ay = fabsf(yc), az = fabsf(zc);
      int face; float fx, fy, fz;
      if (az >= ax && az >= ay) {
        if (zc > 0.f) { face = 0; fx = xc; fy = yc; fz = zc; } else { face = 2; fx = -xc; fy = yc; fz = -zc; }
      } else if (ax >= ay) {
        if (xc > 0.f) { face = 1; fx = -zc; fy = yc; fz = xc; } else { face = 3; fx = zc; fy = yc; fz = -xc; }
      } else {
        if (yc > 0.f) { face = 5; fx = -xc; fy = zc; fz = yc; } else { face = 4; fx = -xc; fy = -zc; fz = -yc; }
      }
      if (!(fz > z_near)) continue;
      float u = fmaf(focal, fx / fz, c);
      float vv = fmaf(focal, fy / fz, c);
      if (!(u >= 0.f && u < fres && vv >= 0.f && vv < fres)) continue;
      int px = (int)floorf(u), py = (int)floorf(vv);
      /* low word: the point index (ties in depth -> lowest index), or — colour-key variant — the packed colour and
       * the low 8 index bits (ties in depth -> lowest packed colour, then lowest index mod 256) */
      const uint32_t low = color_keys ? (((f2u(pts4[i * 4 + 3]) & 0xFFFFFFu) << 8) | (uint32_t)(i & 0xFF)) : (uint32_t)i;
      uint64_t key = ((uint64_t)f2u(fz) << 32) | low;
      uint64_t* cell = zv + ((int64_t)face * res + py) * res + px;
      if (key < *cell) *cell = key;
    }
  }
}

void oracle_splat_keys_cube(const float* pts4, int64_t n, const float* w2c_front, int V, int res, float focal,
                            float z_near, uint64_t* keys) {
  splat_keys_cube_core(pts4, n, w2c_front, V, res, focal, z_near, keys, 0);
}

/* Colour-key variant of the cube splat (EVW_SPLAT_COLOR_KEYS in include/evoworld_b200.h): same geometry, the key carries
 * the colour so that the resolve needs no gather.  Identical panoramas unless two points of DIFFERENT colour have exactly
 * the same float32 depth in the same cell. */
void oracle_splat_keys_cube_colorkey(const float* pts4, int64_t n, const float* w2c_front, int V, int res, float focal,
                                     float z_near, uint64_t* keys) {
  splat_keys_cube_core(pts4, n, w2c_front, V, res, focal, z_near, keys, 1);
}

/* cube->equirect gather through the lookup table (reproject_vggt_open3d_utils.py:603-612):
 * lut [npix] = face<<28 | row<<14 | col, 0xFFFFFFFF = none; out [V,npix,3] */
void oracle_resolve(const uint64_t* keys, const float* pts4, const uint32_t* lut, int V, int res,
                    int64_t npix, uint8_t* out) {
  const int64_t view_cells = (int64_t)6 * res * res;
#pragma omp parallel for
  for (int v = 0; v < V; ++v) {
    for (int64_t p = 0; p < npix; ++p) {
      uint8_t* o = out + ((int64_t)v * npix + p) * 3;
      o[0] = o[1] = o[2] = 0;
      uint32_t e = lut[p];
      if (e == 0xFFFFFFFFu) continue;
      uint32_t face = e >> 28, row = (e >> 14) & 0x3FFFu, col = e & 0x3FFFu;
      uint64_t key = keys[(int64_t)v * view_cells + ((int64_t)face * res + row) * res + col];
      if (key == UINT64_MAX) continue;
      uint32_t rgb = f2u(pts4[(key & 0xFFFFFFFFu) * 4 + 3]);
      o[0] = rgb & 0xFF; o[1] = (rgb >> 8) & 0xFF; o[2] = (rgb >> 16) & 0xFF;
    }
  }
}

/* resolve for colour keys: the colour is bits 8..31 of the key */
void oracle_resolve_colorkey(const uint64_t* keys, const uint32_t* lut, int V, int res, int64_t npix, uint8_t* out) {
  const int64_t view_cells = (int64_t)6 * res * res;
#pragma omp parallel for
  for (int v = 0; v < V; ++v) {
    for (int64_t p = 0; p < npix; ++p) {
      uint8_t* o = out + ((int64_t)v * npix + p) * 3;
      o[0] = o[1] = o[2] = 0;
      uint32_t e = lut[p];
      if (e == 0xFFFFFFFFu) continue;
      uint32_t face = e >> 28, row = (e >> 14) & 0x3FFFu, col = e & 0x3FFFu;
      uint64_t key = keys[(int64_t)v * view_cells + ((int64_t)face * res + row) * res + col];
      if (key == UINT64_MAX) continue;
      uint32_t rgb = (uint32_t)(key >> 8) & 0xFFFFFFu;
      o[0] = rgb & 0xFF; o[1] = (rgb >> 8) & 0xFF; o[2] = (rgb >> 16) & 0xFF;
    }
  }
}

// Host-side feature catalogs and their flattening into integral-image tap tables.
//
// Catalog order must equal the reference's generateFeatures() order because variable indices are
// part of the trained model (cascade.xml "features" list):
//   Haar: src/haarfeatures.cpp:127-251   LBP: src/lbpfeatures.cpp:35-45   HOG: src/HOGfeatures.cpp:67-106
// The shapes are expressed as data (ShapeSpec) rather than as the reference's unrolled branches.
#include <algorithm>
#include <cstring>

#include "cc_internal.h"

namespace cc {

namespace {

// An upright shape: outer rect (x, y, mx*dx, my*dy) weight -1 plus up to two inner rects given in
// units of (dx, dy) relative to (x, y).
struct Inner { int ox, oy, sx, sy; float wt; };
struct UprightSpec { int mx, my; int min_mode; int n_inner; Inner in[2]; };

const UprightSpec kUpright[] = {
    {2, 1, CC_HAAR_BASIC, 1, {{1, 0, 1, 1, 2.f}, {}}},                      // x2
    {1, 2, CC_HAAR_BASIC, 1, {{0, 1, 1, 1, 2.f}, {}}},                      // y2
    {3, 1, CC_HAAR_BASIC, 1, {{1, 0, 1, 1, 2.f}, {}}},                      // x3
    {1, 3, CC_HAAR_BASIC, 1, {{0, 1, 1, 1, 2.f}, {}}},                      // y3
    {4, 1, CC_HAAR_CORE, 1, {{1, 0, 2, 1, 2.f}, {}}},                       // x4
    {1, 4, CC_HAAR_CORE, 1, {{0, 1, 1, 2, 2.f}, {}}},                       // y4
    {2, 2, CC_HAAR_BASIC, 2, {{0, 0, 1, 1, 2.f}, {1, 1, 1, 1, 2.f}}},       // x2_y2 (checkerboard)
    {3, 3, CC_HAAR_CORE, 1, {{1, 1, 1, 1, 9.f}, {}}},                       // centre-surround
};

// A tilted shape.  along_x: the long side runs along the rect's "width" direction (+1,+1) and is
// k*dx; otherwise it runs along the "height" direction (-1,+1) and is k*dy.  inner_ofs / inner_len
// are in units of dx (or dy) along that direction.
struct TiltedSpec { bool along_x; int k; int inner_ofs, inner_len; float wt; };
const TiltedSpec kTilted[] = {
    {true, 2, 0, 1, 2.f},    // tilted x2
    {false, 2, 0, 1, 2.f},   // tilted y2
    {true, 3, 1, 1, 3.f},    // tilted x3
    {false, 3, 1, 1, 3.f},   // tilted y3
    {true, 4, 1, 2, 2.f},    // tilted x4
    {false, 4, 1, 2, 2.f},   // tilted y4
};

}  // namespace

std::vector<HaarFeatureH> make_haar_catalog(int W, int H, int mode) {
  std::vector<HaarFeatureH> out;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y)
      for (int dx = 1; dx <= W; ++dx)
        for (int dy = 1; dy <= H; ++dy) {
          int shape = -1;
          for (const UprightSpec& s : kUpright) {
            ++shape;
            if (mode < s.min_mode) continue;
            if (x + dx * s.mx > W || y + dy * s.my > H) continue;
            HaarFeatureH f;
            std::memset(&f, 0, sizeof(f));
            f.tilted = 0;
            f.r[0] = {x, y, dx * s.mx, dy * s.my};
            f.wt[0] = -1.f;
            for (int k = 0; k < s.n_inner; ++k) {
              const Inner& in = s.in[k];
              f.r[k + 1] = {x + in.ox * dx, y + in.oy * dy, in.sx * dx, in.sy * dy};
              f.wt[k + 1] = in.wt;
            }
            f.shape = shape; f.x = x; f.y = y; f.dx = dx; f.dy = dy;
            out.push_back(f);
          }
          if (mode != CC_HAAR_ALL) continue;
          shape = 7;
          for (const TiltedSpec& s : kTilted) {
            ++shape;
            HaarFeatureH f;
            std::memset(&f, 0, sizeof(f));
            f.tilted = 1;
            f.shape = shape; f.x = x; f.y = y; f.dx = dx; f.dy = dy;
            f.wt[0] = -1.f;
            f.wt[1] = s.wt;
            if (s.along_x) {
              // outer (x, y, k*dx, dy); fits iff x+k*dx<=W, y+k*dx+dy<=H, x-dy>=0
              if (x + s.k * dx > W || y + s.k * dx + dy > H || x - dy < 0) continue;
              f.r[0] = {x, y, s.k * dx, dy};
              f.r[1] = {x + s.inner_ofs * dx, y + s.inner_ofs * dx, s.inner_len * dx, dy};
            } else {
              // outer (x, y, dx, k*dy); fits iff x+dx<=W, y+dx+k*dy<=H, x-k*dy>=0
              if (x + dx > W || y + dx + s.k * dy > H || x - s.k * dy < 0) continue;
              f.r[0] = {x, y, dx, s.k * dy};
              f.r[1] = {x - s.inner_ofs * dy, y + s.inner_ofs * dy, dx, s.inner_len * dy};
            }
            out.push_back(f);
          }
        }
  return out;
}

std::vector<LbpFeatureH> make_lbp_catalog(int W, int H) {
  std::vector<LbpFeatureH> out;
  for (int x = 0; x < W; ++x)
    for (int y = 0; y < H; ++y)
      for (int w = 1; w <= W / 3; ++w)
        for (int h = 1; h <= H / 3; ++h)
          if (x + 3 * w <= W && y + 3 * h <= H) out.push_back({{x, y, w, h}});
  return out;
}

std::vector<HogFeatureH> make_hog_catalog(int W, int H) {
  std::vector<HogFeatureH> out;
  for (int t = 8; t <= W / 2; t += 8) {
    const int shapes[3][2] = {{t, t}, {t, 2 * t}, {2 * t, t}};   // cell (w, h); block = 2x2 cells
    for (const auto& s : shapes)
      for (int x = 0; x <= W - 2 * s[0]; x += 4)
        for (int y = 0; y <= H - 2 * s[1]; y += 4) out.push_back({{x, y, s[0], s[1]}});
  }
  return out;
}

// Corner offsets of the reference (CV_SUM_OFFSETS / CV_TILTED_OFFSETS,
// include/traincascade_features.h:40-63); sum = I[p0] - I[p1] - I[p2] + I[p3].
static void corners(const RectI& r, int S, bool tilted, int p[4]) {
  if (!tilted) {
    p[0] = r.x + S * r.y;
    p[1] = r.x + r.w + S * r.y;
    p[2] = r.x + S * (r.y + r.h);
    p[3] = r.x + r.w + S * (r.y + r.h);
  } else {
    p[0] = r.x + S * r.y;
    p[1] = r.x - r.h + S * (r.y + r.h);
    p[2] = r.x + r.w + S * (r.y + r.w);
    p[3] = r.x + r.w - r.h + S * (r.y + r.w + r.h);
  }
}

// Merge the 4-corner expansions of all rects into unique taps with integer coefficients.  The
// weights are small integers (-1, 2, 3, 9) and every partial sum stays below 2^24 for windows up
// to ~148x148, so sum(coef*I) converted to float equals the reference's float expression
// (include/haarfeatures.h:114-122) bit for bit.
HaarTaps make_haar_taps(const HaarFeatureH& f, int S) {
  HaarTaps t;
  std::memset(&t, 0, sizeof(t));
  int offs[12], coefs[12], n = 0;
  for (int j = 0; j < 3; ++j) {
    if (f.wt[j] == 0.f) break;
    int p[4];
    corners(f.r[j], S, f.tilted != 0, p);
    const int w = (int)f.wt[j];
    const int sign[4] = {+1, -1, -1, +1};
    for (int k = 0; k < 4; ++k) {
      int found = -1;
      for (int q = 0; q < n; ++q)
        if (offs[q] == p[k]) found = q;
      if (found < 0) { offs[n] = p[k]; coefs[n] = 0; found = n++; }
      coefs[found] += sign[k] * w;
    }
  }
  int m = 0;
  for (int q = 0; q < n; ++q)
    if (coefs[q] != 0) { t.w[m] = (uint32_t)(offs[q] & 0xFFFF) | ((uint32_t)(uint8_t)(int8_t)coefs[q] << 16); ++m; }
  t.w[0] |= ((uint32_t)m << 24) | ((uint32_t)(f.tilted ? 1 : 0) << 28);
  return t;
}

LbpTaps make_lbp_taps(const LbpFeatureH& f, int S) {
  LbpTaps t;
  for (int gy = 0; gy < 4; ++gy)
    for (int gx = 0; gx < 4; ++gx) t.off[gy * 4 + gx] = (uint16_t)((f.r.x + gx * f.r.w) + S * (f.r.y + gy * f.r.h));
  return t;
}

HogTaps make_hog_taps(const HogFeatureH& f, int S, int G) {
  HogTaps t;
  const RectI c0 = {f.cell0.x / G, f.cell0.y / G, f.cell0.w / G, f.cell0.h / G};     // exact: all four are multiples of G
  const RectI cells[4] = {{c0.x, c0.y, c0.w, c0.h}, {c0.x + c0.w, c0.y, c0.w, c0.h},
                          {c0.x, c0.y + c0.h, c0.w, c0.h}, {c0.x + c0.w, c0.y + c0.h, c0.w, c0.h}};
  for (int i = 0; i < 4; ++i) {
    int p[4];
    corners(cells[i], S, false, p);
    for (int k = 0; k < 4; ++k) t.off[i][k] = (uint16_t)p[k];
  }
  return t;
}

}  // namespace cc

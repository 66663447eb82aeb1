// Deterministic synthetic KITTI-format sequence generator.
//
// The reference ships no data and there is no network, so every test and benchmark of this repo
// runs on sequences produced here: an HDL-64-like scanner (beam-major, then azimuth -- the file
// order of real KITTI scans) driving through a procedurally generated street.  The generator is a
// data source for BOTH the product and the oracle; it is not part of either.  Layout written by
// vxs_write_sequence() is what the reference's KittiReader expects (README.md:48-55):
//   <dir>/velodyne/%06d.bin   N x 4 float32 (x, y, z, remission)     KittiReader.cpp:145-171
//   <dir>/labels/%06d.label   N x uint32, low 16 bits = class         KittiReader.cpp:173-192
//   <dir>/calib.txt           "Tr: <12 numbers>"                      kitti_utils.cpp:29-58
//   <dir>/poses.txt           12 numbers per line (camera frame)      kitti_utils.cpp:71-105
//
// Everything is a pure function of (params, scan index, ray index): results do not depend on the
// number of threads used.
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <sys/stat.h>

extern "C" {

struct vxs_params {
  uint64_t seed;          // scene seed (default 1234)
  uint32_t n_beams;       // 64
  uint32_t n_azimuth;     // 2000
  float elev_top_deg;     // +2.0
  float elev_bottom_deg;  // -24.8
  float max_return_m;     // 120
  float sensor_height_m;  // 1.73
  float speed_m;          // forward motion per scan (1.0)
  float noise_m;          // +-uniform range noise (0.02)
  float unlabeled_frac;   // 0.03 -> label 0
  float outlier_frac;     // 0.01 -> label 1
  float sway_amp_m;       // 0.3   (y = amp * sin(0.03 t))
  float yaw_amp_rad;      // 0.02  (yaw = amp * sin(0.05 t))
};

void vxs_default_params(vxs_params* p) {
  p->seed = 1234;
  p->n_beams = 64;
  p->n_azimuth = 2000;
  p->elev_top_deg = 2.0f;
  p->elev_bottom_deg = -24.8f;
  p->max_return_m = 120.0f;
  p->sensor_height_m = 1.73f;
  p->speed_m = 1.0f;
  p->noise_m = 0.02f;
  p->unlabeled_frac = 0.03f;
  p->outlier_frac = 0.01f;
  p->sway_amp_m = 0.3f;
  p->yaw_amp_rad = 0.02f;
}

}  // extern "C"

namespace {

inline uint64_t mix64(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {
    s += 0x9E3779B97F4A7C15ull;
    return mix64(s);
  }
  double uni() { return double(next() >> 11) * (1.0 / 9007199254740992.0); }  // [0,1)
};

struct Box {
  double lo[3], hi[3];
  uint32_t label;
  uint32_t instance;
};

const double kCell = 10.0;

// Objects whose x-extent lies inside cell c = [10c, 10c+10).  Pure function of (seed, c).
void cell_objects(uint64_t seed, long c, double ground_z, Box* out, int* n_out) {
  Rng r(mix64(seed ^ (uint64_t(c) * 0xD1B54A32D192ED03ull)));
  int n = 0;
  const double x0 = double(c) * kCell;
  auto add = [&](double xa, double xb, double ya, double yb, double za, double zb, uint32_t label) {
    Box& b = out[n++];
    b.lo[0] = xa; b.hi[0] = xb;
    b.lo[1] = ya; b.hi[1] = yb;
    b.lo[2] = ground_z + za; b.hi[2] = ground_z + zb;
    b.label = label;
    b.instance = uint32_t(r.next() & 0xFFFFu) | 1u;
  };
  for (int side = -1; side <= 1; side += 2) {
    const double s = double(side);
    // building
    if (r.uni() < 0.7) {
      const double a = r.uni() * 2.0, len = 5.0 + r.uni() * 2.9, y = 9.0 + r.uni() * 3.0, h = 4.0 + r.uni() * 5.0;
      const double ya = s > 0 ? y : -(y + 6.0), yb = s > 0 ? y + 6.0 : -y;
      add(x0 + a, x0 + a + len, ya, yb, 0.0, h, 50);
    } else {
      r.uni(); r.uni(); r.uni(); r.uni();
    }
    // parked car
    if (r.uni() < 0.4) {
      const double a = r.uni() * 5.5, y = 4.4 + r.uni() * 0.3;
      const double ya = s > 0 ? y : -(y + 1.8), yb = s > 0 ? y + 1.8 : -y;
      add(x0 + a, x0 + a + 4.2, ya, yb, 0.0, 1.5, 10);
    } else {
      r.uni(); r.uni();
    }
    // pole
    if (r.uni() < 0.5) {
      const double a = r.uni() * 9.5, y = 6.5;
      add(x0 + a, x0 + a + 0.2, s * y - 0.1, s * y + 0.1, 0.0, 5.0, 80);
    } else {
      r.uni();
    }
    // tree: trunk + crown
    if (r.uni() < 0.4) {
      const double a = 1.5 + r.uni() * 6.0, y = 7.6;
      add(x0 + a, x0 + a + 0.4, s * y - 0.2, s * y + 0.2, 0.0, 3.0, 71);
      add(x0 + a - 1.0, x0 + a + 1.4, s * y - 1.2, s * y + 1.2, 3.0, 5.5, 70);
    } else {
      r.uni();
    }
    // person on the sidewalk
    if (r.uni() < 0.1) {
      const double a = r.uni() * 9.0, y = 5.0;
      add(x0 + a, x0 + a + 0.5, s * y - 0.25, s * y + 0.25, 0.0, 1.75, r.uni() < 0.5 ? 30u : 254u);
    } else {
      r.uni(); r.uni();
    }
  }
  // vehicle standing in the opposite lane, labelled as a moving class
  if (r.uni() < 0.15) {
    const double a = r.uni() * 5.0;
    add(x0 + a, x0 + a + 4.4, 2.1, 3.9, 0.0, 1.6, 252);
  }
  *n_out = n;
}

const int kMaxCellObjects = 16;

struct CellCache {
  uint64_t seed;
  double ground_z;
  long base;
  std::vector<int> counts;
  std::vector<Box> boxes;
  CellCache(uint64_t s, double gz, long lo, long hi) : seed(s), ground_z(gz), base(lo) {
    const size_t n = size_t(hi - lo + 1);
    counts.resize(n);
    boxes.resize(n * kMaxCellObjects);
    for (size_t i = 0; i < n; ++i) cell_objects(seed, lo + long(i), ground_z, &boxes[i * kMaxCellObjects], &counts[i]);
  }
  const Box* get(long c, int* n) const {
    const long i = c - base;
    if (i < 0 || size_t(i) >= counts.size()) {
      *n = 0;
      return nullptr;
    }
    *n = counts[size_t(i)];
    return &boxes[size_t(i) * kMaxCellObjects];
  }
};

void pose_of(const vxs_params& p, uint32_t t, double* x, double* y, double* yaw) {
  *x = double(p.speed_m) * double(t);
  *y = double(p.sway_amp_m) * std::sin(0.03 * double(t));
  *yaw = double(p.yaw_amp_rad) * std::sin(0.05 * double(t));
}

uint32_t scan_impl(const vxs_params& p, uint32_t t, float* xyzr, uint32_t* labels, uint32_t cap) {
  double px, py, yaw;
  pose_of(p, t, &px, &py, &yaw);
  const double ground_z = -double(p.sensor_height_m);
  const double maxr = double(p.max_return_m);
  const long clo = long(std::floor((px - maxr) / kCell)) - 1, chi = long(std::floor((px + maxr) / kCell)) + 1;
  const CellCache cache(p.seed, ground_z, clo, chi);
  const double cy = std::cos(yaw), sy = std::sin(yaw);
  const double kPi = 3.14159265358979323846;

  uint32_t n = 0;
  for (uint32_t b = 0; b < p.n_beams; ++b) {
    const double frac = p.n_beams > 1 ? double(b) / double(p.n_beams - 1) : 0.0;
    const double elev = (double(p.elev_top_deg) + frac * (double(p.elev_bottom_deg) - double(p.elev_top_deg))) * kPi / 180.0;
    const double ce = std::cos(elev), se = std::sin(elev);
    for (uint32_t a = 0; a < p.n_azimuth; ++a) {
      const double az = 2.0 * kPi * double(a) / double(p.n_azimuth);
      // direction in the sensor frame, then world frame (yaw about z)
      const double ds[3] = {ce * std::cos(az), ce * std::sin(az), se};
      const double d[3] = {cy * ds[0] - sy * ds[1], sy * ds[0] + cy * ds[1], ds[2]};
      const double o[3] = {px, py, 0.0};

      double best = maxr;
      uint32_t label = 0xFFFFFFFFu, inst = 0;
      if (d[2] < 0.0) {
        const double tg = (ground_z - o[2]) / d[2];
        if (tg < best) {
          best = tg;
          const double yw = std::fabs(o[1] + tg * d[1]);
          label = yw < 4.0 ? 40u : (yw < 6.0 ? 48u : 72u);
          inst = 0;
        }
      }
      // walk x-cells in ray order until the entry distance exceeds the best hit
      long c = long(std::floor(o[0] / kCell));
      const int stepc = d[0] > 0 ? 1 : -1;
      for (;;) {
        int nb;
        const Box* bx = cache.get(c, &nb);
        for (int k = 0; k < nb; ++k) {
          double t0 = 0.0, t1 = best;
          bool ok = true;
          for (int ax = 0; ax < 3 && ok; ++ax) {
            if (std::fabs(d[ax]) < 1e-12) {
              if (o[ax] < bx[k].lo[ax] || o[ax] > bx[k].hi[ax]) ok = false;
            } else {
              double ta = (bx[k].lo[ax] - o[ax]) / d[ax], tb = (bx[k].hi[ax] - o[ax]) / d[ax];
              if (ta > tb) {
                const double tmp = ta;
                ta = tb;
                tb = tmp;
              }
              if (ta > t0) t0 = ta;
              if (tb < t1) t1 = tb;
              if (t0 > t1) ok = false;
            }
          }
          if (ok && t0 > 0.0 && t0 < best) {
            best = t0;
            label = bx[k].label;
            inst = bx[k].instance;
          }
        }
        if (std::fabs(d[0]) < 1e-12) break;
        const double xedge = stepc > 0 ? double(c + 1) * kCell : double(c) * kCell;
        const double texit = (xedge - o[0]) / d[0];
        if (texit >= best) break;
        c += stepc;
      }
      if (label == 0xFFFFFFFFu) continue;  // no return within max_return_m

      Rng r(mix64(p.seed ^ mix64(uint64_t(t) * 0x100000001B3ull + 1000ull) ^ (uint64_t(b) * p.n_azimuth + a)));
      const double range = best + (r.uni() * 2.0 - 1.0) * double(p.noise_m);
      const double rem = r.uni();
      const double u = r.uni();
      if (u < double(p.unlabeled_frac)) {
        label = 0;
      } else if (u < double(p.unlabeled_frac) + double(p.outlier_frac)) {
        label = 1;
      }
      if (n < cap) {
        xyzr[4 * size_t(n) + 0] = float(ds[0] * range);
        xyzr[4 * size_t(n) + 1] = float(ds[1] * range);
        xyzr[4 * size_t(n) + 2] = float(ds[2] * range);
        xyzr[4 * size_t(n) + 3] = float(rem);
        labels[n] = (inst << 16) | (label & 0xFFFFu);
      }
      ++n;
    }
  }
  return n;
}

// KITTI seq-00-like velodyne -> camera calibration (row-major 3x4).
const double kTr[12] = {4.276802385584e-04, -9.999672484946e-01, -8.084491683471e-03, -1.198459927713e-02,
                        -7.210626507497e-03, 8.081198471645e-03, -9.999413164504e-01, -5.403984729748e-02,
                        9.999738645903e-01, 4.859485810390e-04, -7.206933692422e-03, -2.921968648686e-01};

void mat4_mul(const double* a, const double* b, double* c) {  // row-major
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = s;
    }
}

void rigid_inverse(const double* a, double* inv) {  // row-major rigid transform
  for (int i = 0; i < 3; ++i)
    for (int j = 0; j < 3; ++j) inv[i * 4 + j] = a[j * 4 + i];
  for (int i = 0; i < 3; ++i) {
    double s = 0;
    for (int k = 0; k < 3; ++k) s += inv[i * 4 + k] * a[k * 4 + 3];
    inv[i * 4 + 3] = -s;
  }
  inv[12] = inv[13] = inv[14] = 0;
  inv[15] = 1;
}

void velo_pose(const vxs_params& p, uint32_t t, double* V) {
  double x, y, yaw;
  pose_of(p, t, &x, &y, &yaw);
  const double c = std::cos(yaw), s = std::sin(yaw);
  const double m[16] = {c, -s, 0, x, s, c, 0, y, 0, 0, 1, 0, 0, 0, 0, 1};
  std::memcpy(V, m, sizeof(m));
}

}  // namespace

extern "C" {

// Generates scan t.  Returns the number of returns (may exceed cap; only cap are stored).
uint32_t vxs_scan(const vxs_params* p, uint32_t t, float* xyzr, uint32_t* labels, uint32_t cap) {
  return scan_impl(*p, t, xyzr, labels, cap);
}

// Intended velodyne pose of scan t (row-major 4x4, double).
void vxs_velo_pose(const vxs_params* p, uint32_t t, double* V16) { velo_pose(*p, t, V16); }

// Tr (velodyne -> camera), row-major 3x4.
void vxs_calib_tr(double* tr12) { std::memcpy(tr12, kTr, sizeof(kTr)); }

// Generates scans t0 .. t0+n-1 straight into caller memory (e.g. pinned host buffers): scan s goes to
// xyzr + 4*s*stride_points and labels + s*stride_points; counts[s] receives its number of returns.
int vxs_generate_many(const vxs_params* p, uint32_t t0, uint32_t n, float* xyzr, uint32_t* labels, uint64_t stride_points,
                      uint32_t* counts, int n_threads) {
  if (n_threads < 1) n_threads = 1;
  const uint32_t cap = uint32_t(stride_points < (uint64_t)p->n_beams * p->n_azimuth ? stride_points : (uint64_t)p->n_beams * p->n_azimuth);
  auto work = [&](int tid) {
    for (uint32_t s = uint32_t(tid); s < n; s += uint32_t(n_threads)) {
      const uint32_t got = scan_impl(*p, t0 + s, xyzr + 4 * size_t(s) * stride_points, labels + size_t(s) * stride_points, cap);
      counts[s] = got < cap ? got : cap;
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
  for (auto& x : th) x.join();
  return 0;
}

// Writes a KITTI-format sequence directory.  Returns 0 on success.
int vxs_write_sequence(const vxs_params* p, const char* dir, uint32_t n_scans, int write_labels, int n_threads) {
  const std::string base(dir);
  ::mkdir(base.c_str(), 0777);
  ::mkdir((base + "/velodyne").c_str(), 0777);
  if (write_labels) ::mkdir((base + "/labels").c_str(), 0777);

  double Tr[16], TrInv[16];
  std::memcpy(Tr, kTr, sizeof(kTr));
  Tr[12] = Tr[13] = Tr[14] = 0;
  Tr[15] = 1;
  rigid_inverse(Tr, TrInv);
  {
    FILE* f = std::fopen((base + "/calib.txt").c_str(), "w");
    if (!f) return -1;
    for (int k = 0; k < 4; ++k) std::fprintf(f, "P%d: 1 0 0 0 0 1 0 0 0 0 1 0\n", k);
    std::fprintf(f, "Tr:");
    for (int i = 0; i < 12; ++i) std::fprintf(f, " %.12e", kTr[i]);
    std::fprintf(f, "\n");
    std::fclose(f);
  }
  {
    FILE* f = std::fopen((base + "/poses.txt").c_str(), "w");
    if (!f) return -1;
    for (uint32_t t = 0; t < n_scans; ++t) {
      double V[16], tmp[16], P[16];
      velo_pose(*p, t, V);
      mat4_mul(Tr, V, tmp);
      mat4_mul(tmp, TrInv, P);
      for (int i = 0; i < 12; ++i) std::fprintf(f, i ? " %.9g" : "%.9g", P[i]);
      std::fprintf(f, "\n");
    }
    std::fclose(f);
  }
  const uint32_t cap = p->n_beams * p->n_azimuth;
  if (n_threads < 1) n_threads = 1;
  std::vector<int> rc(size_t(n_threads), 0);
  auto work = [&](int tid) {
    std::vector<float> pts(size_t(cap) * 4);
    std::vector<uint32_t> lab(cap);
    for (uint32_t t = uint32_t(tid); t < n_scans; t += uint32_t(n_threads)) {
      const uint32_t n = scan_impl(*p, t, pts.data(), lab.data(), cap);
      char name[64];
      std::snprintf(name, sizeof(name), "/velodyne/%06u.bin", t);
      FILE* f = std::fopen((base + name).c_str(), "wb");
      if (!f) {
        rc[size_t(tid)] = -1;
        return;
      }
      std::fwrite(pts.data(), sizeof(float) * 4, n, f);
      std::fclose(f);
      if (write_labels) {
        std::snprintf(name, sizeof(name), "/labels/%06u.label", t);
        f = std::fopen((base + name).c_str(), "wb");
        if (!f) {
          rc[size_t(tid)] = -1;
          return;
        }
        std::fwrite(lab.data(), sizeof(uint32_t), n, f);
        std::fclose(f);
      }
    }
  };
  std::vector<std::thread> th;
  for (int i = 0; i < n_threads; ++i) th.emplace_back(work, i);
  for (auto& x : th) x.join();
  for (int v : rc)
    if (v) return v;
  return 0;
}

}  // extern "C"

// Exact fp32/fp64 arithmetic of the reference's visibility ray (VoxelGrid::occludedBy,
// /root/reference/src/data/VoxelGrid.cpp:235-316; voxel2position / position2voxel,
// VoxelGrid.h:93-102), usable from device code and -- for the CPU model of the wave algorithm in
// tests/model -- from host code.  Every operation is individually rounded: on the device through
// the __f*_rn / __d*_rn intrinsics (which the compiler never contracts into FMAs), on the host by
// building with -ffp-contract=off on a target without FMA instructions.
//
// Reference build: x86-64 SSE2, -O3, no fast-math (CMakeLists.txt:10-11,25-26): no FMA exists in
// that binary, division and conversions are IEEE round-to-nearest, float -> int32 is cvttss2si
// (truncation; NaN / out-of-range -> INT_MIN).
#ifndef VOXELIZER_B200_VX_DDA_H_
#define VOXELIZER_B200_VX_DDA_H_

#include <stdint.h>

#if defined(__CUDACC__)
#define VX_HD __host__ __device__ __forceinline__
#else
#define VX_HD inline
#endif

#if defined(__CUDA_ARCH__)
#define VX_FADD(a, b) __fadd_rn((a), (b))
#define VX_FSUB(a, b) __fsub_rn((a), (b))
#define VX_FMUL(a, b) __fmul_rn((a), (b))
#define VX_FDIV(a, b) __fdiv_rn((a), (b))
#define VX_DADD(a, b) __dadd_rn((a), (b))
#define VX_DMUL(a, b) __dmul_rn((a), (b))
#define VX_DDIV(a, b) __ddiv_rn((a), (b))
#define VX_D2F(a) __double2float_rn(a)
#define VX_I2F(a) __int2float_rn(a)
#else
#define VX_FADD(a, b) ((a) + (b))
#define VX_FSUB(a, b) ((a) - (b))
#define VX_FMUL(a, b) ((a) * (b))
#define VX_FDIV(a, b) ((a) / (b))
#define VX_DADD(a, b) ((a) + (b))
#define VX_DMUL(a, b) ((a) * (b))
#define VX_DDIV(a, b) ((a) / (b))
#define VX_D2F(a) ((float)(a))
#define VX_I2F(a) ((float)(a))
#endif

struct VxGridGeom {
  int32_t nx, ny, nz;
  float res;
  float off[3];
};

// float -> int32 with x86 cvttss2si semantics.
VX_HD int32_t vx_trunc_i32(float f) {
  if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
  return (int32_t)f;
}

// VoxelGrid.h:93-97: float(double(float(off + float(i*res))) + 0.5*double(res)).
// The reference adds in double and narrows.  Both operands are floats (0.5*res is exact), and a
// double holds 53 >= 2*24 + 2 bits, so rounding the double sum to float gives the correctly rounded
// float sum (double rounding is innocuous for + - * / at that width): one fp32 add is bit-identical.
VX_HD float vx_voxel_centre(float off, float res, int32_t i) {
  const float a = VX_FMUL(VX_I2F(i), res);
  const float b = VX_FADD(off, a);
  return VX_FADD(b, VX_FMUL(0.5f, res));
}

// Per-axis traversal constants (VoxelGrid.cpp:249-270).
struct VxAxis {
  float tmax;    // NextCrossingT
  float tdelta;  // DeltaT
  int32_t step;  // +-1
  int32_t out;   // leaving index: 0 for negative steps (sic), size for positive steps
  int32_t end;   // end-voxel coordinate: truncation toward zero (VoxelGrid.h:99-102)
};

// NextCrossingT = float(+-half / double(dir)) with half = 0.5 * double(res) (VoxelGrid.cpp:257,264): by
// the same argument the double division narrowed to float equals the fp32 division (0.5f*res) / dir,
// and because halving is exact that is 0.5f * DeltaT unless DeltaT is subnormal (then divide again).
VX_HD VxAxis vx_axis_setup(float off, float res, int32_t size, int32_t start, float endpoint) {
  VxAxis a;
  const float dir = VX_FSUB(endpoint, vx_voxel_centre(off, res, start));
  if (dir < 0.0f) {
    a.tdelta = VX_FDIV(-res, dir);
    a.step = -1;
    a.out = 0;
  } else {
    a.tdelta = VX_FDIV(res, dir);
    a.step = 1;
    a.out = size;
  }
  a.tmax = VX_FMUL(0.5f, a.tdelta);
  if (a.tdelta < 4.0e-38f) a.tmax = dir < 0.0f ? VX_FDIV(VX_FMUL(-0.5f, res), dir) : VX_FDIV(VX_FMUL(0.5f, res), dir);
  a.end = vx_trunc_i32(VX_FDIV(VX_FSUB(endpoint, off), res));
  return a;
}

// VoxelGrid.cpp:299-301: bits = (T0<T1)<<2 | (T0<T2)<<1 | (T1<T2); axis = {2,1,2,1,2,2,0,0}[bits]
VX_HD int32_t vx_step_axis(float t0, float t1, float t2) {
  const uint32_t bits = ((uint32_t)(t0 < t1) << 2) | ((uint32_t)(t0 < t2) << 1) | (uint32_t)(t1 < t2);
  return (int32_t)((0x00221212u >> (bits * 4u)) & 3u);
}

// One ray of the reference traversal.  Usage:
//   VxRay r; r.begin(geom, i, j, k, e);
//   for (;;) { if (!r.inside(geom)) break;  /* look at (r.i, r.j, r.k) */
//              if (occupied) { hit; break; }  /* record the cell */
//              if (!r.advance()) break; }
// advance() returns false when the reference loop would `break` after the step (VoxelGrid.cpp:307-308).
// Kept in scalars and stepped without branches: a ray step is the unit of work of the whole path.
struct VxRay {
  int32_t i, j, k;
  float tx, ty, tz;     // NextCrossingT
  float dx, dy, dz;     // DeltaT
  int32_t sx, sy, sz;   // Step
  int32_t ox, oy, oz;   // Out
  int32_t ex, ey, ez;   // end voxel

  VX_HD void begin(const VxGridGeom& g, int32_t si, int32_t sj, int32_t sk, float e0, float e1, float e2) {
    i = si;
    j = sj;
    k = sk;
    const VxAxis ax = vx_axis_setup(g.off[0], g.res, g.nx, si, e0);
    const VxAxis ay = vx_axis_setup(g.off[1], g.res, g.ny, sj, e1);
    const VxAxis az = vx_axis_setup(g.off[2], g.res, g.nz, sk, e2);
    tx = ax.tmax; dx = ax.tdelta; sx = ax.step; ox = ax.out; ex = ax.end;
    ty = ay.tmax; dy = ay.tdelta; sy = ay.step; oy = ay.out; ey = ay.end;
    tz = az.tmax; dz = az.tdelta; sz = az.step; oz = az.out; ez = az.end;
  }

  VX_HD bool inside(const VxGridGeom& g) const {
    return (uint32_t)i < (uint32_t)g.nx && (uint32_t)j < (uint32_t)g.ny && (uint32_t)k < (uint32_t)g.nz;
  }

  // axis the NEXT advance() will move along
  VX_HD int32_t next_axis() const { return vx_step_axis(tx, ty, tz); }

  VX_HD bool advance() {
    const int32_t a = next_axis();
    const bool a0 = a == 0, a1 = a == 1, a2 = a == 2;
    const float nx_ = VX_FADD(tx, dx), ny_ = VX_FADD(ty, dy), nz_ = VX_FADD(tz, dz);
    i += a0 ? sx : 0;
    j += a1 ? sy : 0;
    k += a2 ? sz : 0;
    tx = a0 ? nx_ : tx;
    ty = a1 ? ny_ : ty;
    tz = a2 ? nz_ : tz;
    const int32_t pos = a0 ? i : (a1 ? j : k);
    const int32_t out = a0 ? ox : (a1 ? oy : oz);
    // pos == out: the reference's break after the step.  pos < 0 (a negative step out of layer 0, which
    // `out == 0` does not catch) ends the reference loop at the top of its next iteration (VoxelGrid.cpp:286).
    // Either way nothing more is traversed and the ray misses: after `true` the position is inside the grid.
    if (pos == out || pos < 0) return false;
    if (i == ex && j == ey && k == ez) return false;
    return true;
  }
};

// Visit schedule shared by updateOcclusions / updateInvalid (VoxelGrid.cpp:104-146, 179-221):
// shells o = 0 .. min(Sx, ceil(Sy/2)) - 1, each with three faces
//   face 0: i = Sx-1-o, all j (outer) x all k (inner)
//   face 1: j = o,        i = 0 .. Sx-o-2 (outer) x all k
//   face 2: j = Sy-1-o,   same
VX_HD int32_t vx_num_shells(int32_t nx, int32_t ny) {
  const int32_t half = (ny + 1) / 2;  // ceil(0.5 * Sy)
  return nx < half ? nx : half;
}

#endif  // VOXELIZER_B200_VX_DDA_H_

// fp32 pose algebra of the host layer.  The reference does these with Eigen::Matrix4f
// (KittiReader.cpp:198-202, voxelize_utils.cpp:183, gen_data.cpp:15-17,114); Eigen is not vendored by
// the reference and not installed here, so its arithmetic is restated for the build the reference's
// flags produce (x86-64 SSE2, no FMA; CMakeLists.txt:10-11,25-26):
//   products : per result column ((a0*b0 + a1*b1) + a2*b2) + a3*b3       (lazy coefficient product)
//   inverse  : Eigen's SSE 2x2-block cofactor routine, one scalar 1/det   (LU/arch/Inverse_SSE.h)
//   norm     : sqrtf(x*x + (y*y + z*z))                                   (redux_novec_unroller)
// All matrices are column-major float[16], the storage order of Eigen::Matrix4f.
// Must be compiled without FMA contraction (-ffp-contract=off).
#ifndef VOXELIZER_B200_HOST_POSE_MATH_H_
#define VOXELIZER_B200_HOST_POSE_MATH_H_

#include <cmath>
#include <cstring>

namespace vxb {

struct Matrix4f {
  float m[16];
  float& operator()(int r, int c) { return m[c * 4 + r]; }
  const float& operator()(int r, int c) const { return m[c * 4 + r]; }
  const float* data() const { return m; }
  float* data() { return m; }
  static Matrix4f Identity() {
    Matrix4f r;
    for (int i = 0; i < 16; ++i) r.m[i] = 0.0f;
    r.m[0] = r.m[5] = r.m[10] = r.m[15] = 1.0f;
    return r;
  }
  Matrix4f inverse() const;
};

inline Matrix4f operator*(const Matrix4f& a, const Matrix4f& b) {
  Matrix4f r;
  for (int col = 0; col < 4; ++col) {
    const float b0 = b.m[col * 4 + 0], b1 = b.m[col * 4 + 1], b2 = b.m[col * 4 + 2], b3 = b.m[col * 4 + 3];
    for (int row = 0; row < 4; ++row) {
      float acc = a.m[row] * b0;
      acc = a.m[4 + row] * b1 + acc;
      acc = a.m[8 + row] * b2 + acc;
      acc = a.m[12 + row] * b3 + acc;
      r.m[col * 4 + row] = acc;
    }
  }
  return r;
}

namespace detail {
// the four SSE lanes of the routine, as plain floats
struct Quad {
  float a, b, c, d;
};
inline Quad operator*(Quad x, Quad y) { return {x.a * y.a, x.b * y.b, x.c * y.c, x.d * y.d}; }
inline Quad operator-(Quad x, Quad y) { return {x.a - y.a, x.b - y.b, x.c - y.c, x.d - y.d}; }
inline Quad operator+(Quad x, Quad y) { return {x.a + y.a, x.b + y.b, x.c + y.c, x.d + y.d}; }
inline Quad splat(float s) { return {s, s, s, s}; }
// 2x2 blocks are stored (m00, m10, m01, m11); adj(X)*Y, X*adj(Y) and det as the routine forms them
inline Quad adj_mul(Quad x, Quad y) {  // X# * Y
  return Quad{x.d, x.d, x.a, x.a} * y - Quad{x.b, x.b, x.c, x.c} * Quad{y.c, y.d, y.a, y.b};
}
inline float det(Quad x) { return x.d * x.a - x.b * x.c; }
}  // namespace detail

inline Matrix4f Matrix4f::inverse() const {
  using namespace detail;
  const Quad A{m[0], m[1], m[4], m[5]}, B{m[2], m[3], m[6], m[7]};
  const Quad C{m[8], m[9], m[12], m[13]}, D{m[10], m[11], m[14], m[15]};
  const Quad AB = adj_mul(A, B), DC = adj_mul(D, C);
  const float dA = det(A), dB = det(B), dC = det(C), dD = det(D);

  // trace(A# B D# C)
  const Quad t = Quad{DC.a, DC.c, DC.b, DC.d} * AB;
  const float tr = (t.a + t.c) + (t.b + t.d);

  Quad iD = Quad{C.a, C.a, C.c, C.c} * Quad{AB.a, AB.b, AB.a, AB.b} + Quad{C.b, C.b, C.d, C.d} * Quad{AB.c, AB.d, AB.c, AB.d};
  Quad iA = Quad{B.a, B.a, B.c, B.c} * Quad{DC.a, DC.b, DC.a, DC.b} + Quad{B.b, B.b, B.d, B.d} * Quad{DC.c, DC.d, DC.c, DC.d};
  iD = D * splat(dA) - iD;
  iA = A * splat(dD) - iA;

  const float determinant = (dA * dD + dB * dC) - tr;
  const float rd = 1.0f / determinant;

  Quad iB = D * Quad{AB.d, AB.a, AB.d, AB.a} - Quad{D.b, D.a, D.d, D.c} * Quad{AB.c, AB.b, AB.c, AB.b};
  Quad iC = A * Quad{DC.d, DC.a, DC.d, DC.a} - Quad{A.b, A.a, A.d, A.c} * Quad{DC.c, DC.b, DC.c, DC.b};
  iB = C * splat(dB) - iB;
  iC = B * splat(dC) - iC;

  const Quad sgn{rd, -rd, -rd, rd};
  iA = sgn * iA;
  iB = sgn * iB;
  iC = sgn * iC;
  iD = sgn * iD;

  Matrix4f r;
  const float out[16] = {iA.d, iA.b, iB.d, iB.b, iA.c, iA.a, iB.c, iB.a, iC.d, iC.b, iD.d, iD.b, iC.c, iC.a, iD.c, iD.a};
  std::memcpy(r.m, out, sizeof(out));
  return r;
}

inline float norm3(float x, float y, float z) {
  const float xx = x * x, yy = y * y, zz = z * z;
  const float t = yy + zz;
  return std::sqrt(xx + t);
}

// gen_data.cpp:15-17
inline float poseDistance(const Matrix4f& A, const Matrix4f& B) {
  return norm3(A.m[12] - B.m[12], A.m[13] - B.m[13], A.m[14] - B.m[14]);
}

}  // namespace vxb

#endif

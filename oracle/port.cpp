// ORACLE / TEST INFRASTRUCTURE ONLY.
//
// CPU restatement ("port") of the reference's gen_data hot path, written from scratch with flat
// arrays.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
// may load this; the product (voxelizer_b200/) never does and has no CPU fallback.
//
// PINNING: every function below is checked bit-for-bit (tests/test_oracle_vs_ref.py) against the
// reference's own translation units compiled unmodified into oracle/_ref/ (oracle/Makefile), on
// seeded synthetic sequences and adversarial fixtures.  What remains UNPINNED is the arithmetic of
// the two third-party libraries the reference does not vendor and this image lacks:
//   * Eigen (>= 3.2, unpinned; README.md:9): Matrix4f inverse / products / Vector3f::norm.  Restated
//     here (and in oracle/shim/eigen3) for an x86-64 SSE2, no-FMA build -- see mat_* below.
//   * Boost (>= 1.54, unpinned; README.md:10): lexical_cast decimal -> float (strtof here) and
//     char_separator tokenisation.
// The reference itself has no tests, golden vectors or fixtures.
//
// Each function cites the reference lines it follows (paths relative to /root/reference).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#include <dirent.h>
#include <sys/stat.h>

namespace orc {

// ------------------------------------------------------------------------------------------------
// fp32 linear algebra, column-major 4x4 (restating Eigen on x86-64/SSE2 without FMA)
// ------------------------------------------------------------------------------------------------
struct Mat4 {
  float m[16];
  float at(int r, int c) const { return m[c * 4 + r]; }
};

static Mat4 mat_identity() {
  Mat4 r;
  for (int i = 0; i < 16; ++i) r.m[i] = (i % 5 == 0) ? 1.0f : 0.0f;
  return r;
}

// Eigen lazy coefficient product: per result column ((a0*b0 + a1*b1) + a2*b2) + a3*b3, each
// operation rounded to fp32 (used at KittiReader.cpp:201, voxelize_utils.cpp:183, gen_data.cpp:114).
static Mat4 mat_mul(const Mat4& a, const Mat4& b) {
  Mat4 r;
  for (int c = 0; c < 4; ++c)
    for (int row = 0; row < 4; ++row) {
      float acc = a.m[row] * b.m[c * 4];
      acc = a.m[4 + row] * b.m[c * 4 + 1] + acc;
      acc = a.m[8 + row] * b.m[c * 4 + 2] + acc;
      acc = a.m[12 + row] * b.m[c * 4 + 3] + acc;
      r.m[c * 4 + row] = acc;
    }
  return r;
}

// 4-lane emulation of the SSE instructions Eigen's 4x4 float inverse is written in.
struct L4 {
  float v[4];
};
static L4 shuf(const L4& a, const L4& b, int imm) {
  L4 r;
  r.v[0] = a.v[imm & 3];
  r.v[1] = a.v[(imm >> 2) & 3];
  r.v[2] = b.v[(imm >> 4) & 3];
  r.v[3] = b.v[(imm >> 6) & 3];
  return r;
}
static L4 lmul(const L4& a, const L4& b) { return {{a.v[0] * b.v[0], a.v[1] * b.v[1], a.v[2] * b.v[2], a.v[3] * b.v[3]}}; }
static L4 lsub(const L4& a, const L4& b) { return {{a.v[0] - b.v[0], a.v[1] - b.v[1], a.v[2] - b.v[2], a.v[3] - b.v[3]}}; }
static L4 ladd(const L4& a, const L4& b) { return {{a.v[0] + b.v[0], a.v[1] + b.v[1], a.v[2] + b.v[2], a.v[3] + b.v[3]}}; }
static L4 movelh(const L4& a, const L4& b) { return {{a.v[0], a.v[1], b.v[0], b.v[1]}}; }
static L4 movehl(const L4& a, const L4& b) { return {{b.v[2], b.v[3], a.v[2], a.v[3]}}; }
static L4 bcast0(const L4& a) { return {{a.v[0], a.v[0], a.v[0], a.v[0]}}; }
static float det2(const L4& x) {  // lane 0 of: t = shuf(x,x,0x5F)*x; t0 - t2
  const L4 t = lmul(shuf(x, x, 0x5F), x);
  return t.v[0] - t.v[2];
}

// Matrix4f::inverse() -- Eigen/src/LU/arch/Inverse_SSE.h: 2x2 block cofactors, one 1/det division.
static Mat4 mat_inverse(const Mat4& in) {
  L4 c0, c1, c2, c3;
  std::memcpy(c0.v, in.m + 0, 16);
  std::memcpy(c1.v, in.m + 4, 16);
  std::memcpy(c2.v, in.m + 8, 16);
  std::memcpy(c3.v, in.m + 12, 16);
  const L4 A = movelh(c0, c1), B = movehl(c1, c0), C = movelh(c2, c3), D = movehl(c3, c2);

  const L4 AB = lsub(lmul(shuf(A, A, 0x0F), B), lmul(shuf(A, A, 0xA5), shuf(B, B, 0x4E)));
  const L4 DC = lsub(lmul(shuf(D, D, 0x0F), C), lmul(shuf(D, D, 0xA5), shuf(C, C, 0x4E)));
  const float dA = det2(A), dB = det2(B), dC = det2(C), dD = det2(D);

  L4 d = lmul(shuf(DC, DC, 0xD8), AB);
  L4 iD = ladd(lmul(shuf(C, C, 0xA0), movelh(AB, AB)), lmul(shuf(C, C, 0xF5), movehl(AB, AB)));
  L4 iA = ladd(lmul(shuf(B, B, 0xA0), movelh(DC, DC)), lmul(shuf(B, B, 0xF5), movehl(DC, DC)));

  d = ladd(d, movehl(d, d));
  const float tr = d.v[0] + d.v[1];
  const float d1 = dA * dD, d2 = dB * dC;

  const L4 vdA = {{dA, dA, dA, dA}}, vdB = {{dB, dB, dB, dB}}, vdC = {{dC, dC, dC, dC}}, vdD = {{dD, dD, dD, dD}};
  iD = lsub(lmul(D, vdA), iD);
  iA = lsub(lmul(A, vdD), iA);

  const float det = (d1 + d2) - tr;
  const float rd = 1.0f / det;

  L4 iB = lsub(lmul(D, shuf(AB, AB, 0x33)), lmul(shuf(D, D, 0xB1), shuf(AB, AB, 0x66)));
  L4 iC = lsub(lmul(A, shuf(DC, DC, 0x33)), lmul(shuf(A, A, 0xB1), shuf(DC, DC, 0x66)));
  const L4 vrd = {{rd, -rd, -rd, rd}};
  iB = lsub(lmul(C, vdB), iB);
  iC = lsub(lmul(B, vdC), iC);
  iA = lmul(vrd, iA);
  iB = lmul(vrd, iB);
  iC = lmul(vrd, iC);
  iD = lmul(vrd, iD);

  Mat4 out;
  const L4 r0 = shuf(iA, iB, 0x77), r1 = shuf(iA, iB, 0x22), r2 = shuf(iC, iD, 0x77), r3 = shuf(iC, iD, 0x22);
  std::memcpy(out.m + 0, r0.v, 16);
  std::memcpy(out.m + 4, r1.v, 16);
  std::memcpy(out.m + 8, r2.v, 16);
  std::memcpy(out.m + 12, r3.v, 16);
  (void)bcast0;
  return out;
}

// Vector3f::norm(): sqrtf(x*x + (y*y + z*z))  (voxelize_utils.cpp:188, gen_data.cpp:16)
static float norm3(float x, float y, float z) {
  const float xx = x * x, yy = y * y, zz = z * z;
  const float t = yy + zz;
  return std::sqrt(xx + t);
}

// ------------------------------------------------------------------------------------------------
// strings (rv/string_utils.h:23-24, string_utils.cpp:10-51)
// ------------------------------------------------------------------------------------------------
// rv::trim's default whitespace literal " \0\t\n\r\x0B" is cut at the embedded NUL when it becomes a
// std::string, so only ' ' is stripped.
static std::string trim_spaces(const std::string& s) {
  size_t b = 0;
  while (b < s.size() && s[b] == ' ') ++b;
  size_t e = s.size();
  while (e > b + 1 && s[e - 1] == ' ') --e;  // the reference's end scan stops at end > beg
  if (b >= s.size()) return std::string();
  return s.substr(b, e - b);
}

// rv::split with keep_empty_tokens: every delimiter closes a token; an empty string has NO token
// (boost token_iterator::initialize is never valid on an empty range).
static std::vector<std::string> split_keep(const std::string& s, const std::string& delims) {
  std::vector<std::string> out;
  if (s.empty()) return out;
  std::string cur;
  for (char ch : s) {
    if (delims.find(ch) != std::string::npos) {
      out.push_back(cur);
      cur.clear();
    } else {
      cur.push_back(ch);
    }
  }
  out.push_back(cur);
  return out;
}

struct BadCast : std::runtime_error {
  BadCast() : std::runtime_error("bad lexical cast: source type value could not be interpreted as target") {}
};

static float to_float(const std::string& s) {
  if (s.empty() || std::isspace(static_cast<unsigned char>(s[0]))) throw BadCast();
  char* end = nullptr;
  const float v = std::strtof(s.c_str(), &end);
  if (end != s.c_str() + s.size()) throw BadCast();
  return v;
}

static uint32_t to_u32(const std::string& s) {
  if (s.empty()) throw BadCast();
  size_t i = 0;
  bool neg = false;
  if (s[0] == '+' || s[0] == '-') {
    neg = s[0] == '-';
    i = 1;
  }
  if (i >= s.size()) throw BadCast();
  uint64_t acc = 0;
  for (; i < s.size(); ++i) {
    if (s[i] < '0' || s[i] > '9') throw BadCast();
    acc = acc * 10 + uint64_t(s[i] - '0');
    if (acc > 0xFFFFFFFFull) throw BadCast();
  }
  return neg ? uint32_t(0u - uint32_t(acc)) : uint32_t(acc);
}

// ------------------------------------------------------------------------------------------------
// Config (voxelize_utils.h:8-28, voxelize_utils.cpp:9-171)
// ------------------------------------------------------------------------------------------------
struct Settings {
  float voxel_size = 0.5f;
  float min_extent[4] = {0, -20, -2, 1};
  float max_extent[4] = {40, 20, 1, 1};
  uint32_t max_scans = 100, prior_scans = 0, past_scans = 10;
  float past_distance = 0.0f, min_range = 2.5f, max_range = 25.0f;
  std::vector<uint32_t> ignore;
  std::map<uint32_t, std::vector<uint32_t>> join;
  uint32_t stride_num = 0;
  float stride_distance = 0.0f;
  bool hidecar = true;
};

template <class T, class F>
static std::vector<T> parse_bracket_list(const std::string& raw, F conv) {  // voxelize_utils.cpp:22-41
  const std::string s = trim_spaces(raw);
  std::vector<T> out;
  if (s.empty() || s[0] != '[' || s[s.size() - 1] != ']') {
    std::fprintf(stderr, "Parser Error: %s is not a valid list token!\n", s.c_str());
    return out;
  }
  for (const std::string& tok : split_keep(s.substr(1, s.size() - 2), ",")) out.push_back(conv(trim_spaces(tok)));
  return out;
}

// the std::string specialisation re-joins nested "[..,..]" pieces (voxelize_utils.cpp:43-72)
static std::vector<std::string> parse_bracket_list_of_strings(const std::string& raw) {
  const std::string s = trim_spaces(raw);
  if (s.empty() || s[0] != '[' || s[s.size() - 1] != ']')
    std::fprintf(stderr, "Parser Error: %s is not a valid list token!\n", s.c_str());
  std::vector<std::string> out;
  const size_t inner = s.size() >= 2 ? s.size() - 2 : 0;
  const std::vector<std::string> parts = split_keep(s.size() >= 1 ? s.substr(1, inner) : std::string(), ",");
  for (size_t i = 0; i < parts.size(); ++i) {
    std::string tok = parts[i];
    const size_t open = tok.find('[');
    if (open != std::string::npos && tok.find(']', open) == std::string::npos) {
      std::string nxt;
      do {
        if (i >= parts.size()) break;
        if (i + 1 >= parts.size()) throw std::out_of_range("unterminated nested list");  // reference: UB
        nxt = parts[i + 1];
        tok += "," + nxt;
        ++i;
      } while (nxt.find(']') == std::string::npos);
    }
    out.push_back(tok);
  }
  return out;
}

static std::vector<std::string> parse_brace_pair(const std::string& raw) {  // voxelize_utils.cpp:9-20
  const std::string s = trim_spaces(raw);
  if (s.empty() || s[0] != '{' || s[s.size() - 1] != '}') {
    std::fprintf(stderr, "Parser Error: %s is not a valid dictionary token!\n", s.c_str());
    return {};
  }
  return split_keep(s.substr(1, s.size() - 2), ":");
}

static Settings load_settings(const std::string& path, std::string* log) {  // voxelize_utils.cpp:74-171
  Settings cfg;
  std::ifstream in(path);
  if (!in.is_open()) return cfg;
  std::string line;
  in.peek();
  while (in.good() && !in.eof()) {
    std::getline(in, line);
    const std::string t = trim_spaces(line);
    if (!t.empty() && t[0] == '#') continue;
    std::vector<std::string> kv = split_keep(line, ":");
    if (kv.size() < 2) continue;
    std::string val = kv[1];
    for (size_t i = 2; i < kv.size(); ++i) val += ":" + kv[i];
    const std::string& key = kv[0];  // compared untrimmed
    if (key == "max scans") cfg.max_scans = to_u32(trim_spaces(val));
    else if (key == "max range") cfg.max_range = to_float(trim_spaces(val));
    else if (key == "voxel size") cfg.voxel_size = to_float(trim_spaces(val));
    else if (key == "min range") cfg.min_range = to_float(trim_spaces(val));
    else if (key == "prior scans") cfg.prior_scans = to_u32(trim_spaces(val));
    else if (key == "past scans") cfg.past_scans = to_u32(trim_spaces(val));
    else if (key == "past distance") cfg.past_scans = uint32_t(to_float(trim_spaces(val)));  // sic, :120-123
    else if (key == "stride num") cfg.stride_num = to_u32(trim_spaces(val));
    else if (key == "stride distance") cfg.stride_distance = to_float(trim_spaces(val));
    else if (key == "min extent" || key == "max extent") {
      const std::vector<float> c = parse_bracket_list<float>(val, to_float);
      if (c.size() < 3) throw std::out_of_range("extent needs three coordinates");  // reference: UB
      float* dst = key == "min extent" ? cfg.min_extent : cfg.max_extent;
      dst[0] = c[0];
      dst[1] = c[1];
      dst[2] = c[2];
      dst[3] = 1.0f;
    } else if (key == "ignore") cfg.ignore = parse_bracket_list<uint32_t>(val, to_u32);
    else if (key == "join") {
      for (const std::string& item : parse_bracket_list_of_strings(trim_spaces(val))) {
        const std::vector<std::string> pr = parse_brace_pair(item);
        if (pr.size() < 2) throw std::out_of_range("join entry needs key and list");  // reference: UB
        cfg.join[to_u32(trim_spaces(pr[0]))] = parse_bracket_list<uint32_t>(trim_spaces(pr[1]), to_u32);
      }
    } else if (log) {
      *log += "unknown parameter: " + key + "\n";
    }
  }
  return cfg;
}

// ------------------------------------------------------------------------------------------------
// Voxel grid (VoxelGrid.h:17-135, VoxelGrid.cpp:5-316)
// ------------------------------------------------------------------------------------------------
// float -> int32 as x86 cvttss2si does it: truncation, NaN / out-of-range -> INT_MIN.
static int32_t trunc_i32(float f) {
  if (!(f > -2147483904.0f && f < 2147483648.0f)) return INT32_MIN;
  return int32_t(f);
}

struct Grid {
  float res = 0;
  int32_t nx = 0, ny = 0, nz = 0;
  float off[3] = {0, 0, 0};
  std::vector<uint32_t> count;                      // per voxel, internal index i + j*nx + k*nx*ny
  std::unordered_map<uint64_t, uint32_t> hist;      // (voxel << 32 | label) -> count
  std::vector<uint32_t> touched;                    // occupied_ list
  std::vector<int32_t> occl, inval, memo;           // occlusions_, invalid_, occludedBy_
  bool occl_valid = false;

  size_t nvox() const { return size_t(nx) * size_t(ny) * size_t(nz); }
  int32_t idx(int32_t i, int32_t j, int32_t k) const { return i + j * nx + k * nx * ny; }

  void setup(float resolution, const float* mn, const float* mx) {  // VoxelGrid.cpp:5-28
    reset();
    res = resolution;
    nx = int32_t(uint32_t(std::ceil((mx[0] - mn[0]) / res)));
    ny = int32_t(uint32_t(std::ceil((mx[1] - mn[1]) / res)));
    nz = int32_t(uint32_t(std::ceil((mx[2] - mn[2]) / res)));
    count.assign(nvox(), 0);
    const uint32_t sz[3] = {uint32_t(nx), uint32_t(ny), uint32_t(nz)};
    for (int a = 0; a < 3; ++a) {
      const float span = mx[a] - mn[a];
      const float cover = float(sz[a]) * res;   // uint32 -> float, float product
      const float slack = cover - span;
      off[a] = float(double(mn[a]) - 0.5 * double(slack));
    }
    occl.assign(nvox(), 0);
    memo.assign(nvox(), 0);
  }

  void reset() {  // VoxelGrid.cpp:30-43
    for (uint32_t v : touched) count[v] = 0;
    hist.clear();
    touched.clear();
  }

  void add_point(const float* p, uint32_t label) {  // VoxelGrid.cpp:45-63
    const float tx = p[0] - off[0], ty = p[1] - off[1], tz = p[2] - off[2];
    const float fi = std::floor(tx / res), fj = std::floor(ty / res), fk = std::floor(tz / res);
    const int32_t i = trunc_i32(fi), j = trunc_i32(fj), k = trunc_i32(fk);
    if (i >= nx || j >= ny || k >= nz) return;
    if (i < 0 || j < 0 || k < 0) return;
    const int32_t g = idx(i, j, k);
    if (g < 0 || size_t(g) >= nvox()) return;
    if (count[size_t(g)] == 0) touched.push_back(uint32_t(g));
    hist[(uint64_t(uint32_t(g)) << 32) | label] += 1;
    count[size_t(g)] += 1;
    occl_valid = false;
  }

  // VoxelGrid.h:93-97: offset + i*res in fp32, "+ 0.5*res" in double, narrowed.
  float centre(int axis, int32_t i) const {
    const float a = float(i) * res;
    const float b = off[axis] + a;
    return float(double(b) + 0.5 * double(res));
  }

  // VoxelGrid.cpp:235-316.  Returns hit voxel index or -1; memoises into `memo`.
  int32_t cast(int32_t i, int32_t j, int32_t k, const float* e, std::vector<uint32_t>& path, std::vector<int32_t>* visited = nullptr) {
    if (visited) visited->clear();  // VoxelGrid.cpp:243
    int32_t pos[3] = {i, j, k};
    const int32_t dim[3] = {nx, ny, nz};
    float tmax[3], tdelta[3];
    int32_t step[3], out[3];
    const double half = 0.5 * double(res);
    for (int a = 0; a < 3; ++a) {
      const float d = e[a] - centre(a, pos[a]);
      if (d < 0) {
        tmax[a] = float(-half / double(d));
        tdelta[a] = -res / d;
        step[a] = -1;
        out[a] = 0;
      } else {
        tmax[a] = float(half / double(d));
        tdelta[a] = res / d;
        step[a] = 1;
        out[a] = dim[a];
      }
    }
    // VoxelGrid.h:99-102: float division, truncation toward zero
    const int32_t end[3] = {trunc_i32((e[0] - off[0]) / res), trunc_i32((e[1] - off[1]) / res), trunc_i32((e[2] - off[2]) / res)};
    static const int32_t pick[8] = {2, 1, 2, 1, 2, 2, 0, 0};
    path.clear();
    for (;;) {
      if (pos[0] < 0 || pos[1] < 0 || pos[2] < 0) break;
      if (pos[0] >= nx || pos[1] >= ny || pos[2] >= nz) break;
      const int32_t v = idx(pos[0], pos[1], pos[2]);
      if (visited) visited->insert(visited->end(), {pos[0], pos[1], pos[2]});  // VoxelGrid.cpp:291
      if (count[size_t(v)] > 0) {
        for (uint32_t t : path) memo[t] = v;
        return v;
      }
      path.push_back(uint32_t(v));
      const int bits = (int(tmax[0] < tmax[1]) << 2) + (int(tmax[0] < tmax[2]) << 1) + int(tmax[1] < tmax[2]);
      const int32_t a = pick[bits];
      pos[a] += step[a];
      tmax[a] += tdelta[a];
      if (pos[a] == out[a]) break;
      if (pos[0] == end[0] && pos[1] == end[1] && pos[2] == end[2]) break;
    }
    for (uint32_t t : path) memo[t] = -1;
    return -1;
  }

  // Visit order shared by updateOcclusions / updateInvalid (VoxelGrid.cpp:104-146, 179-221).
  template <class Visit>
  void sweep(Visit visit) {
    const uint32_t shells = std::min<uint32_t>(uint32_t(nx), uint32_t(std::ceil(0.5 * double(uint32_t(ny)))));
    for (uint32_t o = 0; o < shells; ++o) {
      const int32_t fi = nx - int32_t(o) - 1;
      for (int32_t j = 0; j < ny; ++j)
        for (int32_t k = 0; k < nz; ++k) visit(fi, j, k);
      for (int pass = 0; pass < 2; ++pass) {
        const int32_t fj = pass == 0 ? int32_t(o) : ny - int32_t(o) - 1;
        for (int32_t i = 0; i < nx - int32_t(o) - 1; ++i)
          for (int32_t k = 0; k < nz; ++k) visit(i, fj, k);
      }
    }
  }

  void occlusion_pass() {  // VoxelGrid.cpp:98-171
    std::fill(memo.begin(), memo.end(), -2);
    std::fill(occl.begin(), occl.end(), -1);
    const float origin[3] = {0, 0, 0};
    std::vector<uint32_t> path;
    sweep([&](int32_t i, int32_t j, int32_t k) {
      const int32_t v = idx(i, j, k);
      if (memo[size_t(v)] == -2) memo[size_t(v)] = cast(i, j, k, origin, path);
      occl[size_t(v)] = memo[size_t(v)];
    });
    inval = occl;
    occl_valid = true;
  }

  void invalid_pass(const float* e) {  // VoxelGrid.cpp:173-233
    if (!occl_valid) occlusion_pass();
    std::fill(memo.begin(), memo.end(), -2);
    std::vector<uint32_t> path;
    sweep([&](int32_t i, int32_t j, int32_t k) {
      const int32_t v = idx(i, j, k);
      if (inval[size_t(v)] == -1) return;
      if (memo[size_t(v)] == -2) memo[size_t(v)] = cast(i, j, k, e, path);
      inval[size_t(v)] = std::min(inval[size_t(v)], memo[size_t(v)]);
    });
  }

  // VoxelGrid.cpp:75-96: every voxel that was not occupied when the occlusions were computed takes count and labels of
  // the first occupied voxel above it, if only occluded empty voxels lie between ("find label from above").
  void fill_from_above() {
    if (!occl_valid) occlusion_pass();
    for (int32_t i = 0; i < nx; ++i)
      for (int32_t j = 0; j < ny; ++j)
        for (int32_t k = 0; k < nz; ++k) {
          const int32_t g = idx(i, j, k);
          if (occl[size_t(g)] == g) continue;
          int32_t n = 1;
          while (k + n < nz && occluded(i, j, k + n) && count[size_t(idx(i, j, k + n))] == 0) n += 1;
          if (k + n < nz && count[size_t(idx(i, j, k + n))] > 0) {
            const int32_t src = idx(i, j, k + n);
            if (count[size_t(g)] == 0) touched.push_back(uint32_t(g));
            count[size_t(g)] = count[size_t(src)];
            // labels = labels of the source voxel (a map assignment: what the voxel held before is dropped)
            std::vector<uint64_t> drop;
            for (const auto& kv : hist)
              if (uint32_t(kv.first >> 32) == uint32_t(g)) drop.push_back(kv.first);
            for (uint64_t key : drop) hist.erase(key);
            std::vector<std::pair<uint64_t, uint32_t>> add;
            for (const auto& kv : hist)
              if (uint32_t(kv.first >> 32) == uint32_t(src)) add.emplace_back((uint64_t(uint32_t(g)) << 32) | (kv.first & 0xFFFFFFFFu), kv.second);
            for (const auto& kv : add) hist[kv.first] = kv.second;
          }
        }
  }

  bool occluded(int32_t i, int32_t j, int32_t k) const { return occl[size_t(idx(i, j, k))] > -1; }  // :65
  bool invalid(int32_t i, int32_t j, int32_t k) const {                                              // :69-73
    const int32_t v = idx(i, j, k);
    if (int32_t(inval.size()) <= v) return true;
    return inval[size_t(v)] > -1 && inval[size_t(v)] != v;
  }

  // voxelize_utils.cpp:237-245: ascending labels, strict '>' -> largest count, smallest label on ties.
  std::vector<uint16_t> majority_internal() const {
    std::vector<uint32_t> best_count(nvox(), 0);
    std::vector<uint32_t> best_label(nvox(), 0);
    for (const auto& kv : hist) {
      const size_t v = size_t(kv.first >> 32);
      const uint32_t label = uint32_t(kv.first & 0xFFFFFFFFu), c = kv.second;
      if (c > best_count[v] || (c == best_count[v] && c > 0 && label < best_label[v])) {
        best_count[v] = c;
        best_label[v] = label;
      }
    }
    std::vector<uint16_t> out(nvox());
    for (size_t v = 0; v < nvox(); ++v) out[v] = uint16_t(best_label[v]);
    return out;
  }
};

// voxelize_utils.cpp:205-216: byte n <- elements 8n..8n+7, element 8n in bit 7; length N/8.
template <class T>
static std::vector<uint8_t> pack_msb(const std::vector<T>& v) {
  std::vector<uint8_t> out(v.size() / 8);
  for (size_t b = 0; b < out.size(); ++b) {
    uint8_t x = 0;
    for (int r = 0; r < 8; ++r) x = uint8_t(x | ((v[8 * b + size_t(r)] > 0 ? 1u : 0u) << (7 - r)));
    out[b] = x;
  }
  return out;
}

struct FrameOut {
  std::vector<uint16_t> label;     // output order x, y, z (z fastest)
  std::vector<uint8_t> occluded;   // unpacked 0/1, output order
  std::vector<uint8_t> invalid;
};

// the triple loop of saveVoxelGrid (voxelize_utils.cpp:229-255)
static FrameOut extract(const Grid& g) {
  FrameOut o;
  const std::vector<uint16_t> maj = g.majority_internal();
  o.label.reserve(g.nvox());
  o.occluded.reserve(g.nvox());
  o.invalid.reserve(g.nvox());
  for (int32_t x = 0; x < g.nx; ++x)
    for (int32_t y = 0; y < g.ny; ++y)
      for (int32_t z = 0; z < g.nz; ++z) {
        o.label.push_back(maj[size_t(g.idx(x, y, z))]);
        o.occluded.push_back(g.occl.empty() ? 0 : uint8_t(g.occluded(x, y, z)));
        o.invalid.push_back(uint8_t(g.invalid(x, y, z)));
      }
  return o;
}

static void write_file(const std::string& path, const void* data, size_t bytes) {
  std::ofstream out(path.c_str(), std::ios::binary);
  out.write(static_cast<const char*>(data), std::streamsize(bytes));
}

static void save_grid(const Grid& g, const std::string& dir, const std::string& base, const std::string& mode) {
  const FrameOut o = extract(g);  // voxelize_utils.cpp:218-296
  if (mode == "target") {
    write_file(dir + "/" + base + ".label", o.label.data(), o.label.size() * 2);
    const std::vector<uint8_t> po = pack_msb(o.occluded), pi = pack_msb(o.invalid);
    write_file(dir + "/" + base + ".occluded", po.data(), po.size());
    write_file(dir + "/" + base + ".invalid", pi.data(), pi.size());
  } else {
    const std::vector<uint8_t> pb = pack_msb(o.label);
    write_file(dir + "/" + base + ".bin", pb.data(), pb.size());
  }
}

// ------------------------------------------------------------------------------------------------
// fillVoxelGrid (voxelize_utils.cpp:173-203)
// ------------------------------------------------------------------------------------------------
struct Scan {
  Mat4 pose;
  std::vector<float> xyzr;       // n x 4
  std::vector<uint32_t> labels;  // n (already & 0xFFFF when it comes from the reader)
  size_t size() const { return labels.size(); }
};

static void fill(const Mat4& anchor, const std::vector<const Scan*>& scans, Grid& g, const Settings& cfg) {
  std::map<uint32_t, uint32_t> member_to_key;
  for (const auto& j : cfg.join)
    for (uint32_t member : j.second) member_to_key[member] = j.first;
  const Mat4 anchor_inv = mat_inverse(anchor);
  for (const Scan* s : scans) {
    const Mat4 ap = mat_mul(anchor_inv, s->pose);
    const size_t n = s->xyzr.size() / 4;
    for (size_t i = 0; i < n; ++i) {
      const float x = s->xyzr[4 * i], y = s->xyzr[4 * i + 1], z = s->xyzr[4 * i + 2];
      const float range = norm3(x, y, z);
      if (range < cfg.min_range || range > cfg.max_range) continue;
      if (cfg.hidecar && double(x) < 3.0 && double(x) > -2.0 && double(std::fabs(y)) < 2.0) continue;
      float p[4];
      for (int r = 0; r < 4; ++r) {  // ((m0*x + m1*y) + m2*z) + m3*1
        float acc = ap.m[r] * x;
        acc = ap.m[4 + r] * y + acc;
        acc = ap.m[8 + r] * z + acc;
        acc = ap.m[12 + r] * 1.0f + acc;
        p[r] = acc;
      }
      const uint32_t original = s->labels[i];
      uint32_t mapped = original;
      const auto it = member_to_key.find(mapped);
      if (it != member_to_key.end()) mapped = it->second;
      if (std::find(cfg.ignore.begin(), cfg.ignore.end(), mapped) == cfg.ignore.end()) g.add_point(p, original);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// KITTI reader (KittiReader.cpp:12-203, kitti_utils.cpp:29-58,71-105)
// ------------------------------------------------------------------------------------------------
struct Sequence {
  std::vector<std::string> velodyne, labels;
  std::vector<Mat4> poses;
  std::map<uint32_t, Scan> cache;
  int32_t n_prior = 0, n_past = 10;

  static bool is_dir(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0 && S_ISDIR(st.st_mode);
  }
  static bool exists(const std::string& p) {
    struct stat st;
    return ::stat(p.c_str(), &st) == 0;
  }
  static uint64_t file_size(const std::string& p) {
    std::ifstream in(p.c_str());
    in.seekg(0, std::ios::end);
    return uint64_t(in.tellg());
  }

  void open(const std::string& dir) {
    velodyne.clear();
    const std::string vdir = dir + "/velodyne";
    if (!is_dir(vdir)) throw std::runtime_error("Missing velodyne files.");
    std::vector<std::string> names;
    if (DIR* d = ::opendir(vdir.c_str())) {
      while (struct dirent* e = ::readdir(d)) {
        const std::string n = e->d_name;
        if (n.empty() || n[0] == '.') continue;
        struct stat st;
        if (::stat((vdir + "/" + n).c_str(), &st) == 0 && S_ISREG(st.st_mode)) names.push_back(n);
      }
      ::closedir(d);
    }
    std::sort(names.begin(), names.end());
    for (const std::string& n : names) velodyne.push_back(vdir + "/" + n);

    if (!exists(dir + "/calib.txt")) throw std::runtime_error("Missing calibration file: " + dir + "/calib.txt");
    const Mat4 Tr = load_calib_tr(dir + "/calib.txt");
    poses = load_poses(dir + "/poses.txt");
    const Mat4 Tr_inv = mat_inverse(Tr);
    for (Mat4& p : poses) p = mat_mul(mat_mul(Tr_inv, p), Tr);  // KittiReader.cpp:198-202

    const std::string ldir = dir + "/labels";
    if (!is_dir(ldir)) ::mkdir(ldir.c_str(), 0777);
    for (size_t i = 0; i < velodyne.size(); ++i) {
      const uint32_t npts = uint32_t(file_size(velodyne[i]) / 16);
      std::string base = names[i];
      const size_t dot = base.find('.');
      if (dot != std::string::npos) base = base.substr(0, dot);
      const std::string lf = ldir + "/" + base + ".label";
      if (!exists(lf)) {  // side effect kept: KittiReader.cpp:43-50
        const std::vector<uint32_t> zeros(npts, 0);
        write_file(lf, zeros.data(), size_t(npts) * 4);
      }
      labels.push_back(lf);
    }
  }

  static Mat4 load_calib_tr(const std::string& path) {  // kitti_utils.cpp:29-58 + operator[] :23-27
    std::ifstream in(path.c_str());
    if (!in.is_open()) throw std::runtime_error("Unable to open calibration file: " + path);
    std::map<std::string, Mat4> found;
    std::string line;
    in.peek();
    while (!in.eof()) {
      std::getline(in, line);
      const std::vector<std::string> kv = split_keep(line, ":");
      if (kv.size() == 2) {
        const std::vector<std::string> e = split_keep(trim_spaces(kv[1]), " ");
        if (e.size() == 12) {
          Mat4 m = mat_identity();
          for (int i = 0; i < 12; ++i) m.m[(i % 4) * 4 + i / 4] = to_float(trim_spaces(e[size_t(i)]));
          found[trim_spaces(kv[0])] = m;
        }
      }
      in.peek();
    }
    if (found.find("Tr") == found.end()) throw std::runtime_error("Unknown name for calibration matrix.");
    return found["Tr"];
  }

  static std::vector<Mat4> load_poses(const std::string& path) {  // kitti_utils.cpp:71-105
    std::ifstream in(path.c_str());
    if (!in.is_open()) throw std::runtime_error("Unable to open pose file :" + path);
    std::vector<Mat4> out;
    std::string line;
    in.peek();
    while (in.good()) {
      std::getline(in, line);
      const std::vector<std::string> e = split_keep(line, " ");
      if (e.size() >= 12) {
        Mat4 m = mat_identity();
        for (int i = 0; i < 12; ++i) m.m[(i % 4) * 4 + i / 4] = to_float(trim_spaces(e[size_t(i)]));
        out.push_back(m);
      }
      in.peek();
    }
    return out;
  }

  static void load_scan(const std::string& bin, const std::string& lab, Scan& s) {  // KittiReader.cpp:145-192
    s.xyzr.clear();
    s.labels.clear();
    size_t npts = 0;
    {
      std::ifstream in(bin.c_str(), std::ios::binary);
      if (in.is_open()) {
        in.seekg(0, std::ios::end);
        npts = size_t(uint32_t(uint64_t(in.tellg()) / 16));
        in.seekg(0, std::ios::beg);
        s.xyzr.resize(npts * 4);
        in.read(reinterpret_cast<char*>(s.xyzr.data()), std::streamsize(npts * 16));
      }
    }
    {
      std::ifstream in(lab.c_str(), std::ios::binary);
      if (in.is_open()) {
        in.seekg(0, std::ios::end);
        const size_t nl = size_t(uint32_t(uint64_t(in.tellg()) / 4));
        in.seekg(0, std::ios::beg);
        s.labels.resize(nl);
        in.read(reinterpret_cast<char*>(s.labels.data()), std::streamsize(nl * 4));
        for (uint32_t& v : s.labels) v &= 0xFFFFu;
      } else {
        std::fprintf(stderr, "Unable to open label file. \n");
      }
    }
    if (npts != s.labels.size()) throw std::runtime_error("Inconsistent number of labels.");
  }

  // KittiReader.cpp:56-143 (window + cache eviction).  Returns number of scans read from disk.
  uint32_t window(int32_t idx, std::vector<const Scan*>& prior, std::vector<const Scan*>& past) {
    prior.clear();
    past.clear();
    const int32_t n = int32_t(velodyne.size());
    if (idx >= n || idx < 0) return 0;
    std::vector<uint32_t> keep;
    uint32_t reads = 0;
    auto fetch = [&](int32_t t) -> const Scan* {
      keep.push_back(uint32_t(t));
      auto it = cache.find(uint32_t(t));
      if (it == cache.end()) {
        ++reads;
        Scan& s = cache[uint32_t(t)];
        load_scan(velodyne[size_t(t)], labels[size_t(t)], s);
        s.pose = poses.at(size_t(t));  // reference: unchecked operator[]
        return &s;
      }
      return &it->second;
    };
    for (int32_t t = std::max<int32_t>(0, idx - n_prior); t <= idx; ++t) prior.push_back(fetch(t));
    for (int32_t t = std::min<int32_t>(n - 1, idx + 1); t < std::min<int32_t>(n, idx + n_past + 1); ++t) past.push_back(fetch(t));
    for (auto it = cache.begin(); it != cache.end();) {
      if (std::find(keep.begin(), keep.end(), it->first) == keep.end()) it = cache.erase(it);
      else ++it;
    }
    return reads;
  }
};

// gen_data.cpp:129-141: stride walk. Returns false when the loop must stop.
static bool advance(uint32_t& current, const std::vector<Mat4>& poses, const Settings& cfg) {
  float distance = 0.0f;
  uint32_t count = 0;
  auto more = [&]() { return count < cfg.stride_num || distance < cfg.stride_distance || count == 0; };
  while (more() && size_t(current) + 1 < poses.size()) {
    const Mat4& a = poses[current];
    const Mat4& b = poses[current + 1];
    distance += norm3(a.m[12] - b.m[12], a.m[13] - b.m[13], a.m[14] - b.m[14]);
    count += 1;
    current += 1;
  }
  if (more() && size_t(current) + 1 >= poses.size()) return false;
  return true;
}

struct FrameStats {
  double fill_s = 0, occl_s = 0, invalid_s = 0, save_s = 0;
  uint32_t frames = 0, skipped = 0;
};

static double now_s() {
  timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return double(ts.tv_sec) + 1e-9 * double(ts.tv_nsec);
}

// gen_data.cpp:19-145.  first/last bound the TARGET LIST positions processed (for sampling the CPU
// baseline and for sharding tests); the walk itself is always the reference's.
static int generate(const std::string& cfg_path, const std::string& seq_dir, const std::string& out_dir,
                    int64_t first_frame, int64_t max_frames, FrameStats* stats, bool quiet) {
  std::string log;
  const Settings cfg = load_settings(cfg_path, &log);
  if (!quiet && !log.empty()) std::fputs(log.c_str(), stdout);
  if (!Sequence::is_dir(out_dir)) {
    std::string cur;
    for (size_t i = 0; i <= out_dir.size(); ++i) {
      if ((i == out_dir.size() || out_dir[i] == '/') && !cur.empty() && !Sequence::is_dir(cur)) {
        if (::mkdir(cur.c_str(), 0777) != 0) throw std::runtime_error("Unable to create output directory.");
      }
      if (i < out_dir.size()) cur.push_back(out_dir[i]);
    }
  }
  Sequence seq;
  seq.open(seq_dir);
  seq.n_prior = int32_t(cfg.prior_scans);
  seq.n_past = int32_t(cfg.past_scans);
  Grid prior_grid, past_grid;
  prior_grid.setup(cfg.voxel_size, cfg.min_extent, cfg.max_extent);
  past_grid.setup(cfg.voxel_size, cfg.min_extent, cfg.max_extent);

  uint32_t current = 0;
  int64_t frame_no = 0;
  while (current < seq.velodyne.size()) {
    const bool selected = frame_no >= first_frame && (max_frames < 0 || frame_no < first_frame + max_frames);
    if (max_frames >= 0 && frame_no >= first_frame + max_frames) break;
    if (selected) {
      char name[32];
      std::snprintf(name, sizeof(name), "%06u", current);
      std::vector<const Scan*> prior, past;
      const uint32_t reads = seq.window(int32_t(current), prior, past);
      if (!quiet) std::printf("%u point clouds read.\n", reads);
      uint32_t labelled = 0, total = 0;
      for (const Scan* s : past) {
        total += uint32_t(s->labels.size());
        for (uint32_t l : s->labels) labelled += l > 0 ? 1u : 0u;
      }
      const float percentage = 100.0f * float(labelled) / float(total);  // 0/0 -> NaN -> skipped
      prior_grid.reset();
      past_grid.reset();
      if (percentage > 0.0f) {
        const Mat4 anchor = prior.back()->pose;
        double t0 = now_s();
        fill(anchor, prior, prior_grid, cfg);
        fill(anchor, prior, past_grid, cfg);
        fill(anchor, past, past_grid, cfg);
        double t1 = now_s();
        prior_grid.occlusion_pass();
        past_grid.occlusion_pass();
        double t2 = now_s();
        const Mat4 anchor_inv = mat_inverse(anchor);
        for (const Scan* s : past) {
          const Mat4 rel = mat_mul(anchor_inv, s->pose);
          const float e[3] = {rel.m[12], rel.m[13], rel.m[14]};
          past_grid.invalid_pass(e);
        }
        double t3 = now_s();
        save_grid(prior_grid, out_dir, name, "input");
        save_grid(past_grid, out_dir, name, "target");
        double t4 = now_s();
        if (stats) {
          stats->fill_s += t1 - t0;
          stats->occl_s += t2 - t1;
          stats->invalid_s += t3 - t2;
          stats->save_s += t4 - t3;
          stats->frames += 1;
        }
      } else {
        if (!quiet) std::printf("skipped.\n");
        if (stats) stats->skipped += 1;
      }
    }
    ++frame_no;
    if (!advance(current, seq.poses, cfg)) break;
  }
  return 0;
}

}  // namespace orc

// ------------------------------------------------------------------------------------------------
// C ABI (loaded by oracle/oracle.py with ctypes)
// ------------------------------------------------------------------------------------------------
static std::string g_err;

extern "C" {

const char* orc_last_error() { return g_err.c_str(); }

struct orc_config_pod {
  float voxelSize;
  float minExtent[4];
  float maxExtent[4];
  uint32_t maxNumScans, priorScans, pastScans;
  float pastDistance, minRange, maxRange;
  uint32_t stride_num;
  float stride_distance;
  int32_t hidecar;
  uint32_t n_filtered;
  uint32_t n_join_pairs;
};

void* orc_config_load(const char* path) {
  try {
    return new orc::Settings(orc::load_settings(path, nullptr));
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_config_free(void* c) { delete static_cast<orc::Settings*>(c); }
void orc_config_get(void* c, orc_config_pod* o) {
  const orc::Settings& s = *static_cast<orc::Settings*>(c);
  o->voxelSize = s.voxel_size;
  for (int i = 0; i < 4; ++i) {
    o->minExtent[i] = s.min_extent[i];
    o->maxExtent[i] = s.max_extent[i];
  }
  o->maxNumScans = s.max_scans;
  o->priorScans = s.prior_scans;
  o->pastScans = s.past_scans;
  o->pastDistance = s.past_distance;
  o->minRange = s.min_range;
  o->maxRange = s.max_range;
  o->stride_num = s.stride_num;
  o->stride_distance = s.stride_distance;
  o->hidecar = s.hidecar ? 1 : 0;
  o->n_filtered = uint32_t(s.ignore.size());
  uint32_t n = 0;
  for (const auto& j : s.join) n += uint32_t(j.second.size());
  o->n_join_pairs = n;
}
void orc_config_filtered(void* c, uint32_t* out) {
  const orc::Settings& s = *static_cast<orc::Settings*>(c);
  for (size_t i = 0; i < s.ignore.size(); ++i) out[i] = s.ignore[i];
}
void orc_config_joins(void* c, uint32_t* out_pairs) {
  const orc::Settings& s = *static_cast<orc::Settings*>(c);
  size_t n = 0;
  for (const auto& j : s.join)
    for (uint32_t v : j.second) {
      out_pairs[2 * n] = j.first;
      out_pairs[2 * n + 1] = v;
      ++n;
    }
}

void* orc_grid_new(float resolution, const float* min4, const float* max4) {
  orc::Grid* g = new orc::Grid;
  g->setup(resolution, min4, max4);
  return g;
}
void orc_grid_free(void* g) { delete static_cast<orc::Grid*>(g); }
void orc_grid_dims(void* gp, uint32_t* size3, float* offset4, float* resolution) {
  const orc::Grid& g = *static_cast<orc::Grid*>(gp);
  size3[0] = uint32_t(g.nx);
  size3[1] = uint32_t(g.ny);
  size3[2] = uint32_t(g.nz);
  offset4[0] = g.off[0];
  offset4[1] = g.off[1];
  offset4[2] = g.off[2];
  offset4[3] = 1.0f;
  *resolution = g.res;
}
void orc_grid_clear(void* gp) { static_cast<orc::Grid*>(gp)->reset(); }
void orc_grid_insert(void* gp, const float* p4, const uint32_t* labels, uint64_t n) {
  orc::Grid& g = *static_cast<orc::Grid*>(gp);
  for (uint64_t i = 0; i < n; ++i) g.add_point(p4 + 4 * i, labels[i]);
}
void orc_fill(void* gp, void* cfgp, const float* anchor16, uint32_t nscans, const float* const* xyzr,
              const uint32_t* const* labels, const uint32_t* counts, const float* poses16) {
  orc::Grid& g = *static_cast<orc::Grid*>(gp);
  std::vector<orc::Scan> scans(nscans);
  std::vector<const orc::Scan*> ptrs;
  for (uint32_t s = 0; s < nscans; ++s) {
    std::memcpy(scans[s].pose.m, poses16 + 16 * size_t(s), 64);
    scans[s].xyzr.assign(xyzr[s], xyzr[s] + 4 * size_t(counts[s]));
    scans[s].labels.assign(labels[s], labels[s] + counts[s]);
    ptrs.push_back(&scans[s]);
  }
  orc::Mat4 anchor;
  std::memcpy(anchor.m, anchor16, 64);
  orc::fill(anchor, ptrs, g, *static_cast<orc::Settings*>(cfgp));
}
void orc_update_occlusions(void* gp) { static_cast<orc::Grid*>(gp)->occlusion_pass(); }
void orc_insert_occlusion_labels(void* gp) { static_cast<orc::Grid*>(gp)->fill_from_above(); }
// VoxelGrid::occludedBy as a public call: returns the occluder (internal index) or -1; visited = (i,j,k) triples
int32_t orc_occluded_by(void* gp, int32_t i, int32_t j, int32_t k, const float* e3, int32_t* visited, uint32_t capacity, uint32_t* n_visited) {
  orc::Grid& g = *static_cast<orc::Grid*>(gp);
  std::vector<uint32_t> path;
  std::vector<int32_t> seen;
  const int32_t r = g.cast(i, j, k, e3, path, &seen);
  const uint32_t n = uint32_t(seen.size() / 3);
  if (n_visited) *n_visited = n;
  for (uint32_t q = 0; q < n && q < capacity; ++q)
    for (int a = 0; a < 3; ++a) visited[3 * q + uint32_t(a)] = seen[3 * q + uint32_t(a)];
  return r;
}
void orc_update_invalid(void* gp, const float* e3) { static_cast<orc::Grid*>(gp)->invalid_pass(e3); }
uint64_t orc_grid_dump(void* gp, int what, int32_t* out) {
  const orc::Grid& g = *static_cast<orc::Grid*>(gp);
  const std::vector<int32_t>& v = what == 0 ? g.occl : (what == 1 ? g.inval : g.memo);
  if (out) std::memcpy(out, v.data(), v.size() * sizeof(int32_t));
  return v.size();
}
void orc_grid_counts(void* gp, uint32_t* out) {
  const orc::Grid& g = *static_cast<orc::Grid*>(gp);
  std::memcpy(out, g.count.data(), g.count.size() * 4);
}
uint64_t orc_grid_hist(void* gp, uint32_t* out) {
  const orc::Grid& g = *static_cast<orc::Grid*>(gp);
  std::vector<std::pair<uint64_t, uint32_t>> items(g.hist.begin(), g.hist.end());
  std::sort(items.begin(), items.end());
  if (out)
    for (size_t n = 0; n < items.size(); ++n) {
      out[3 * n] = uint32_t(items[n].first >> 32);
      out[3 * n + 1] = uint32_t(items[n].first & 0xFFFFFFFFu);
      out[3 * n + 2] = items[n].second;
    }
  return items.size();
}
void orc_grid_flags(void* gp, uint8_t* occluded, uint8_t* invalid) {
  const orc::FrameOut o = orc::extract(*static_cast<orc::Grid*>(gp));
  std::memcpy(occluded, o.occluded.data(), o.occluded.size());
  std::memcpy(invalid, o.invalid.data(), o.invalid.size());
}
// majority label per voxel in OUTPUT order (x, y, z fastest)
void orc_grid_labels(void* gp, uint16_t* out) {
  const orc::FrameOut o = orc::extract(*static_cast<orc::Grid*>(gp));
  std::memcpy(out, o.label.data(), o.label.size() * 2);
}
void orc_grid_save(void* gp, const char* dir, const char* basename, const char* mode) {
  orc::save_grid(*static_cast<orc::Grid*>(gp), dir, basename, mode);
}

void orc_mat_inverse(const float* a16, float* out16) {
  orc::Mat4 a;
  std::memcpy(a.m, a16, 64);
  const orc::Mat4 r = orc::mat_inverse(a);
  std::memcpy(out16, r.m, 64);
}
void orc_mat_mul(const float* a16, const float* b16, float* out16) {
  orc::Mat4 a, b;
  std::memcpy(a.m, a16, 64);
  std::memcpy(b.m, b16, 64);
  const orc::Mat4 r = orc::mat_mul(a, b);
  std::memcpy(out16, r.m, 64);
}
void orc_endpoint(const float* anchor16, const float* pose16, float* e3) {
  orc::Mat4 a, p;
  std::memcpy(a.m, anchor16, 64);
  std::memcpy(p.m, pose16, 64);
  const orc::Mat4 r = orc::mat_mul(orc::mat_inverse(a), p);
  e3[0] = r.m[12];
  e3[1] = r.m[13];
  e3[2] = r.m[14];
}

void* orc_reader_open(const char* dir) {
  try {
    orc::Sequence* s = new orc::Sequence;
    s->open(dir);
    return s;
  } catch (const std::exception& e) {
    g_err = e.what();
    return nullptr;
  }
}
void orc_reader_free(void* r) { delete static_cast<orc::Sequence*>(r); }
uint32_t orc_reader_count(void* r) { return uint32_t(static_cast<orc::Sequence*>(r)->velodyne.size()); }
uint32_t orc_reader_num_poses(void* r) { return uint32_t(static_cast<orc::Sequence*>(r)->poses.size()); }
void orc_reader_poses(void* r, float* out16n) {
  const orc::Sequence& s = *static_cast<orc::Sequence*>(r);
  for (size_t i = 0; i < s.poses.size(); ++i) std::memcpy(out16n + 16 * i, s.poses[i].m, 64);
}

// Whole program.  stats6 (optional): fill_s, occl_s, invalid_s, save_s, frames, skipped.
int orc_gen_data(const char* cfg, const char* seq_dir, const char* out_dir, int64_t first_frame, int64_t max_frames,
                 double* stats6, int quiet) {
  try {
    orc::FrameStats st;
    const int rc = orc::generate(cfg, seq_dir, out_dir, first_frame, max_frames, &st, quiet != 0);
    if (stats6) {
      stats6[0] = st.fill_s;
      stats6[1] = st.occl_s;
      stats6[2] = st.invalid_s;
      stats6[3] = st.save_s;
      stats6[4] = double(st.frames);
      stats6[5] = double(st.skipped);
    }
    return rc;
  } catch (const std::exception& e) {
    g_err = e.what();
    return -1;
  }
}

}  // extern "C"

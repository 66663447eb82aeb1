// K4: the visibility passes -- VoxelGrid::updateOcclusions (VoxelGrid.cpp:98-171), one
// VoxelGrid::updateInvalid per past scan (VoxelGrid.cpp:173-233) and the memoised ray they share,
// VoxelGrid::occludedBy (VoxelGrid.cpp:235-316) -- reproduced bit for bit.
//
// The reference is order dependent: a ray writes its result into occludedBy_ for every cell it
// crossed, unconditionally, and a cell visited later reads whatever the LATEST ray that crossed it
// left there (casting its own ray only if nothing has crossed it yet).  Casting one independent ray
// per voxel flips ~60 k output bits per frame.  This kernel keeps the reference's visit order and
// parallelises inside it (tests/model/wave_model.cpp is the CPU model of exactly this formulation):
//
//   * LIVE cells.  A visit matters only for a free cell that, in an invalid pass, has not been killed
//     yet (invalid_ != -1).  Occupied cells answer occludedBy() with their own index at once, are
//     never recorded by any ray (a ray stops in front of them), so they stay "occluded" and "not
//     invalid" for good and are left out.  One bitmap (`alive`) holds the live set; a typical
//     invalid pass touches < 20 % of the 2.6 M visits of the reference's sweep.
//   * a WAVE = a run of whole columns of one face of one shell holding <= kMaxLive live cells,
//     compacted into shared memory.  One CTA owns one frame and runs its waves in visit order:
//       A  snapshot: owner[r] = id of the latest ray that crossed live cell r (global memo, tagged
//          with the pass number so that "fill occludedBy_ with -2" costs nothing); cells nothing has
//          crossed are CANDIDATES.
//       B  which candidates cast?  c casts iff no CASTING candidate visited earlier crosses c with
//          the in-face part of its ray (rays are monotone per axis: they never re-enter the face).
//          Counters: cnt[c] = undecided earlier candidates crossing c.  killed -> no cast (releases
//          its successors); cnt == 0 -> casts (kills its successors).  Monotone, so any thread
//          interleaving gives the reference's answer; chains resolve in as many sweeps as they are long
//          (one barrier per sweep; a thread remembers in registers which candidates its candidate's ray
//          crosses, so a sweep does not walk the ray again).
//       C  casters get ray ids in visit order and trace: RED.MAX(memo[cell], id) = "latest ray wins";
//          hit[id] is one bit set when the ray ends on an occupied cell.  Rays are pulled from a
//          queue by the lanes of few warps and refilled as they end, so that a long ray does not
//          idle a whole warp (most rays are ~25 cells, a few are > 200).
//       D  apply the visit: hit[owner] decides -- occlusions_ (last visit wins) in pass 0, kill in
//          an invalid pass (invalid_ = min(...) can only go to -1 once).
//   * only booleans are observable (isOccluded / isInvalid, VoxelGrid.cpp:65-73).
//
// Latency- and issue-bound, not HBM-bound.  The occupancy test of a ray step reads a shared-memory
// bitmap of 4x4x1-voxel blocks (5 % of the steps need the exact bit from global memory); leaving the
// grid is one zero-field test on the packed position (trace_warp).  Global traffic: one memo read per live visit,
// one fire-and-forget RED per ray step, one hit-bit gather per live visit.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <type_traits>

#include "vx_launch.h"

// Checked build (-DVX_CHECKED, tools/build_variants.py): every index into shared memory and into the per-frame global
// arrays is tested and a violation traps (compute-sanitizer is closed on the B200 pool, so the bounds checks are our own).
#ifdef VX_CHECKED
#define VX_ASSERT(cond)                                                                               \
  do {                                                                                                \
    if (!(cond)) {                                                                                    \
      printf("VX_ASSERT failed: %s (vx_visibility.cu:%d, block %d thread %d)\n", #cond, __LINE__, (int)blockIdx.x, (int)threadIdx.x); \
      __trap();                                                                                       \
    }                                                                                                 \
  } while (0)
#else
#define VX_ASSERT(cond) ((void)0)
#endif

namespace {

// Shape of one frame's CTA.  Two instantiations (bottom of the file):
//   * THROUGHPUT: 128 threads, waves of <= 1280 live cells, 7 CTAs per SM -- a batch of >= ~1000 frames fills the device with
//     frames in flight;
//   * LATENCY: 512 threads, waves of <= 8192 live cells (a whole face of a 256 x 256 x 32 grid), one CTA per SM -- for batches
//     that cannot fill the device (a GUI rebuild, the 512 x 512 x 64 stress grid, 1001-scan windows): a frame is a chain of
//     waves, each a few block-wide phases, so fewer and larger waves on more threads cut the time of ONE frame ~3x.
template <int THREADS, int MAXLIVE, int MAXCHUNKS, int CTAS_PER_SM, int COARSE_WORDS = 4096>
struct VisCfg {
  static constexpr int kThreads = THREADS;
  static constexpr int kWarps = THREADS / 32;
  static constexpr int kMaxLive = MAXLIVE;      // live cells per wave
  static constexpr int kMaxChunks = MAXCHUNKS;  // chunk = <= 32 consecutive z of one column
  static constexpr int kChunksPerThread = (MAXCHUNKS + THREADS - 1) / THREADS;
  static constexpr int kLivePerThread = (MAXLIVE + THREADS - 1) / THREADS;
  static constexpr int kCtasPerSm = CTAS_PER_SM;
  static constexpr uint32_t kCoarseWords = COARSE_WORDS;  // 32 block bits each; a power of two
  // casters per tracing warp when a wave's rays are dealt out: a full warp where frames in flight fill the device (most lanes
  // per instruction), half a warp where one frame has an SM to itself (more warps, shorter critical path)
  static constexpr uint32_t kRaysPerTracer = THREADS >= 256 ? 16u : 32u;
  static_assert(THREADS % 32 == 0 && MAXLIVE % 32 == 0 && MAXLIVE / 32 <= THREADS, "bitsets are cleared by one thread per word");
  static_assert(MAXLIVE < 32768 && MAXCHUNKS <= 2048, "live ranks and chunk << 5 | bit are kept in 16 bits (bit 15 is a flag)");
};
constexpr int kAxisTable = 560;   // nx + ny + nz entries of the per-pass ray set-up table (256 + 256 + 32 = 544)
constexpr uint32_t kSet = 0x80000000u;
constexpr uint32_t kEpochShift = 25;
constexpr uint32_t kIdMask = (1u << kEpochShift) - 1u;
constexpr uint32_t kMaxEpoch = 127;       // epochs 1..127 (0 = never written)
constexpr uint32_t kRefillLanes = 8;      // idle lanes that trigger a refill of the tracer warp

template <class C>
struct VisShared {
  uint32_t coarse[C::kCoarseWords];    // bit per (2^s x 2^s x 1) block: any voxel occupied
  uint32_t owner[C::kMaxLive];         // per live cell: 0 unset | kSet + ray id; cnt[] of candidates during B
  uint16_t rcell[C::kMaxLive];         // live rank -> chunk << 5 | bit
  uint16_t cand[C::kMaxLive];          // live ranks of the candidates, then of the casters, in visit order
  uint32_t chmask[C::kMaxChunks];      // live mask of a chunk
  uint16_t chbase[C::kMaxChunks];      // live rank of its first live cell
  uint32_t killed[C::kMaxLive / 32];   // per live rank: crossed by a decided caster
  uint32_t decided[C::kMaxLive / 32];  // per candidate slot
  uint32_t castbit[C::kMaxLive / 32];  // per candidate slot
  uint32_t has_succ[C::kMaxLive / 32]; // per candidate slot: its in-face ray crosses a later candidate
  float axis[kAxisTable];           // per pass: +-DeltaT of a ray starting at x = i | y = j | z = k (sign = Step)
  uint32_t warp_sums[C::kWarps];
  uint32_t scan_total;
  uint32_t ncols_ok;
  uint32_t any_edge;
  uint32_t undecided;
  uint32_t queue;
  uint32_t scratch_id;
  uint32_t maxlen;  // (statistics) longest ray of the wave, in steps
  // pass 0 of a frame with insertOcclusionLabels (null otherwise; kept here, not in registers: a rare path):
  uint32_t* hitcell;   // ray id -> L of the cell that stopped it
  uint32_t* occluder;  // per free cell the reference-internal index of its occluder (last visit wins)
};

struct Wave {
  int32_t face;     // 0: i fixed, outer j; 1/2: j fixed, outer i
  int32_t fixed;    // the fixed coordinate
  int32_t a_begin;  // first outer coordinate
  int32_t ncols;
};

struct Stats {
  unsigned long long rays = 0, steps = 0, waves = 0, resolve_waves = 0, rounds = 0;
  unsigned long long cyc[4] = {0, 0, 0, 0};
  unsigned long long longest = 0, iters = 0, attns = 0;  // developer statistics (VX_TRACE_STATS): sum over waves of the longest ray,
                                                         // step-loop iterations and attention events as seen by thread 0
};

struct Frame {
  VxGridGeom g;
  int32_t cpc;       // chunks per column = ceil(nz / 32)
  // coarse occupancy: bit k & 31 of word (bi << csI | bj << csJ | k >> 5), bi = i >> cshift, bj = j >> cshift: all strides are
  // powers of two, so a ray step finds its word with shifts and masks only.  A grid too large for the bitmap gets
  // cshift = 11 (one block) with every bit set: "look at the exact bit" everywhere.
  uint32_t cshift, csI, csJ, cmask;
  const uint32_t* __restrict__ occ;
  uint32_t* live;    // [words] free and (invalid pass) not killed
  uint32_t* memo;    // [nvox] epoch << 25 | id of the latest ray of that pass
  uint32_t* hitbits; // [visits / 32 + 1] ray id -> ended on an occupied cell (this pass)
  uint32_t* occl;
  float e0, e1, e2;
  int32_t ex, ey, ez;  // end voxel of the pass (VoxelGrid.h:99-102: truncation)
  uint32_t pend;       // ... as a packed position; one that cannot be packed (outside the grid) cannot be reached either
  uint32_t epoch;
  uint32_t nvox, hit_words;  // (bounds of memo / hitbits, for the checked build)
  bool inv;
  bool tables;        // sm.axis holds this pass's ray set-up
};

__device__ __forceinline__ uint32_t lin(const VxGridGeom& g, int32_t i, int32_t j, int32_t k) {
  return ((uint32_t)i * (uint32_t)g.ny + (uint32_t)j) * (uint32_t)g.nz + (uint32_t)k;
}

// bits [L0, L0 + len) of a bitmap, len <= 32
__device__ __forceinline__ uint32_t extract_bits(const uint32_t* bits, uint32_t L0, uint32_t len) {
  const uint32_t w = L0 >> 5, s = L0 & 31u;
  const uint32_t lo = __ldcg(bits + w);
  const uint32_t hi = (s + len > 32u) ? __ldcg(bits + w + 1) : 0u;
  const uint32_t v = __funnelshift_r(lo, hi, s);
  return len >= 32u ? v : (v & ((1u << len) - 1u));
}

// exclusive prefix sum of one value per thread over the CTA; total in sm.scan_total (valid after return)
template <class C>
__device__ __forceinline__ uint32_t block_scan(VisShared<C>& sm, uint32_t v) {
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  uint32_t x = v;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
    if (lane >= (uint32_t)off) x += y;
  }
  __syncthreads();  // previous users of warp_sums / scan_total are done
  if (lane == 31u) sm.warp_sums[warp] = x;
  __syncthreads();
  uint32_t before = 0, total = 0;
#pragma unroll
  for (int w = 0; w < C::kWarps; ++w) {
    const uint32_t s = sm.warp_sums[w];
    if ((uint32_t)w < warp) before += s;
    total += s;
  }
  if (threadIdx.x == 0) sm.scan_total = total;
  return before + x - v;
}

__device__ __forceinline__ void wave_cell(const Frame& f, const Wave& w, uint32_t rc, int32_t& i, int32_t& j, int32_t& k) {
  const uint32_t c = rc >> 5, bit = rc & 31u;
  uint32_t col, part;
  if (f.cpc == 1) {
    col = c;
    part = 0;
  } else {
    col = c / (uint32_t)f.cpc;
    part = c - col * (uint32_t)f.cpc;
  }
  k = (int32_t)(part * 32u + bit);
  const int32_t outer = w.a_begin + (int32_t)col;
  if (w.face == 0) {
    i = w.fixed;
    j = outer;
  } else {
    i = outer;
    j = w.fixed;
  }
}

// live rank of (i,j,k) inside the wave, or -1
template <class C>
__device__ __forceinline__ int32_t live_rank(const Frame& f, const Wave& w, const VisShared<C>& sm, int32_t i, int32_t j, int32_t k) {
  const int32_t fx = w.face == 0 ? i : j;
  const int32_t col = (w.face == 0 ? j : i) - w.a_begin;
  if (fx != w.fixed || (uint32_t)col >= (uint32_t)w.ncols) return -1;
  const uint32_t c = (uint32_t)col * (uint32_t)f.cpc + ((uint32_t)k >> 5);
  const uint32_t bit = (uint32_t)k & 31u;
  VX_ASSERT(c < (uint32_t)C::kMaxChunks);
  const uint32_t m = sm.chmask[c];
  if (!((m >> bit) & 1u)) return -1;
  return (int32_t)((uint32_t)sm.chbase[c] + (uint32_t)__popc(m & ((1u << bit) - 1u)));
}

__device__ __forceinline__ uint32_t coarse_word(const Frame& f, uint32_t i, uint32_t j, uint32_t k) {
  return ((i >> f.cshift) << f.csI) | ((j >> f.cshift) << f.csJ) | (k >> 5);
}

template <class C>
__device__ __forceinline__ bool coarse_set(const Frame& f, const VisShared<C>& sm, int32_t i, int32_t j, int32_t k) {
  return (sm.coarse[coarse_word(f, (uint32_t)i, (uint32_t)j, (uint32_t)k)] >> ((uint32_t)k & 31u)) & 1u;
}

__device__ __forceinline__ bool occupied_exact(const Frame& f, int32_t i, int32_t j, int32_t k) {
  const uint32_t L = lin(f.g, i, j, k);
  return (__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u;
}

template <class C>
__device__ __forceinline__ bool occupied(const Frame& f, const VisShared<C>& sm, int32_t i, int32_t j, int32_t k) {
  return coarse_set(f, sm, i, j, k) && occupied_exact(f, i, j, k);
}

// Ray set-up (VoxelGrid.cpp:249-270).  DeltaT, Step and NextCrossingT of an axis depend only on that axis'
// start coordinate and the pass's endpoint, so a pass tabulates them once (nx + ny + nz divisions) and a ray
// starts with three shared-memory loads instead of three IEEE divisions.
template <class C>
__device__ __forceinline__ void ray_begin(const Frame& f, const VisShared<C>& sm, VxRay& r, int32_t i, int32_t j, int32_t k) {
  if (!f.tables) {
    r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
    return;
  }
  const float vx = sm.axis[i], vy = sm.axis[f.g.nx + j], vz = sm.axis[f.g.nx + f.g.ny + k];
  r.i = i;
  r.j = j;
  r.k = k;
  r.dx = fabsf(vx);
  r.dy = fabsf(vy);
  r.dz = fabsf(vz);
  r.tx = VX_FMUL(0.5f, r.dx);
  r.ty = VX_FMUL(0.5f, r.dy);
  r.tz = VX_FMUL(0.5f, r.dz);
  r.sx = signbit(vx) ? -1 : 1;
  r.sy = signbit(vy) ? -1 : 1;
  r.sz = signbit(vz) ? -1 : 1;
  r.ox = r.sx < 0 ? 0 : f.g.nx;
  r.oy = r.sy < 0 ? 0 : f.g.ny;
  r.oz = r.sz < 0 ? 0 : f.g.nz;
  r.ex = f.ex;
  r.ey = f.ey;
  r.ez = f.ez;
}

// Tabulates the pass (all threads).  Leaves f.tables false when the grid is too large for the table or when
// an entry would need the reference's special cases (subnormal DeltaT, a -0.0 direction).
template <class C>
__device__ void build_axis_table(Frame& f, VisShared<C>& sm) {
  const int32_t total = f.g.nx + f.g.ny + f.g.nz;
  f.tables = false;
  if (total > kAxisTable) return;
  bool bad = false;
  for (int32_t x = (int32_t)threadIdx.x; x < total; x += C::kThreads) {
    const int32_t axis = x < f.g.nx ? 0 : (x < f.g.nx + f.g.ny ? 1 : 2);
    const int32_t c = axis == 0 ? x : (axis == 1 ? x - f.g.nx : x - f.g.nx - f.g.ny);
    const float e = axis == 0 ? f.e0 : (axis == 1 ? f.e1 : f.e2);
    const int32_t size = axis == 0 ? f.g.nx : (axis == 1 ? f.g.ny : f.g.nz);
    const float off = axis == 0 ? f.g.off[0] : (axis == 1 ? f.g.off[1] : f.g.off[2]);
    const VxAxis a = vx_axis_setup(off, f.g.res, size, c, e);
    // the table encodes Step in the sign and relies on NextCrossingT == 0.5 * DeltaT
    if (!(a.tdelta >= 4.0e-38f) || a.tmax != VX_FMUL(0.5f, a.tdelta)) bad = true;
    sm.axis[x] = a.step < 0 ? -a.tdelta : a.tdelta;
  }
  f.tables = __syncthreads_or(bad ? 1 : 0) == 0;
}

// The first step of the ray from (i,j,k) is along `normal`: its in-face part is empty.
template <class C>
__device__ __forceinline__ bool normal_first(const Frame& f, const VisShared<C>& sm, int32_t i, int32_t j, int32_t k, int32_t normal) {
  if (f.tables) {
    const float tx = VX_FMUL(0.5f, fabsf(sm.axis[i])), ty = VX_FMUL(0.5f, fabsf(sm.axis[f.g.nx + j]));
    const float tz = VX_FMUL(0.5f, fabsf(sm.axis[f.g.nx + f.g.ny + k]));
    return vx_step_axis(tx, ty, tz) == normal;
  }
  // without the table: certainly along `normal` when its |dir| dominates (NextCrossingT = half / |dir|; the
  // 1e-4 margin is far above the rounding of the division and of the conversion to float)
  const float d0 = fabsf(VX_FSUB(f.e0, vx_voxel_centre(f.g.off[0], f.g.res, i)));
  const float d1 = fabsf(VX_FSUB(f.e1, vx_voxel_centre(f.g.off[1], f.g.res, j)));
  const float d2 = fabsf(VX_FSUB(f.e2, vx_voxel_centre(f.g.off[2], f.g.res, k)));
  const float dn = normal == 0 ? d0 : d1;
  const float other = fmaxf(normal == 0 ? d1 : d0, d2);
  return dn > 1.0001f * other;
}

// In-face part of candidate rc's ray: fn(x) for every live wave cell x visited after rc that it crosses
// before it leaves the face, hits an occupied cell or terminates.
template <class C, class Fn>
__device__ __forceinline__ void prefix_walk(const Frame& f, const Wave& w, const VisShared<C>& sm, uint32_t rc, Fn fn) {
  int32_t i, j, k;
  wave_cell(f, w, sm.rcell[rc], i, j, k);
  VxRay r;
  ray_begin(f, sm, r, i, j, k);
  const int32_t normal = w.face == 0 ? 0 : 1;
  for (;;) {
    if (r.next_axis() == normal) return;
    if (!r.advance()) return;  // a ray that advances stays inside the grid (it breaks on the boundary layer)
    if (occupied(f, sm, r.i, r.j, r.k)) return;
    const int32_t x = live_rank(f, w, sm, r.i, r.j, r.k);
    if (x > (int32_t)rc) fn((uint32_t)x);
  }
}

// Phase C: the lanes of the calling warp pull casters from sm.queue and walk the reference's
// occludedBy() once per ray.  The loop body is the unit of work of the whole path, so the ray lives in
// scalars: the same operations in the same order as VxRay (vx_dda.h), with the linear index kept
// incrementally and the boundary tests of VoxelGrid.cpp:286-287,307 folded into one unsigned compare.
// Position of a ray as one word: i | j << 11 | k << 22 (vx_create refuses grids beyond 2047 x 2047 x 1023).
// A step is one add; a coordinate that leaves the grid on the low side wraps to the field's maximum (and
// borrows from the next field), which the unsigned range test below reads as "outside", like a coordinate
// that reaches the size.
constexpr uint32_t kPosJ = 11, kPosK = 22, kPosMaskIJ = 0x7FFu;

template <class C>
__device__ void trace_warp(const Frame& f, const Wave& w, VisShared<C>& sm, uint32_t n_cast, uint32_t id_base, uint32_t first_grab, Stats& st) {
  const uint32_t lane = threadIdx.x & 31u;
  const int32_t normal = w.face == 0 ? 0 : 1;
  const uint32_t enc_epoch = f.epoch << kEpochShift;
  const int32_t stride_j = f.g.nz, stride_i = f.g.ny * f.g.nz;
  // Leaving the grid (VoxelGrid.cpp:286-287,307) as ONE test on the packed position: OUT holds, per axis, the coordinate
  // a step must not reach -- the size for a positive Step, 0 for a negative one (sic: Out = 0, the ray ends ON layer 0), or
  // the wrapped -1 when the ray starts on layer 0 (it can only step out).  No coordinate of a ray that is still inside
  // equals its OUT value, so "some field of P ^ OUT is zero" is exactly "the step left the grid".
  constexpr uint32_t kFieldLow = 1u | (1u << kPosJ) | (1u << kPosK), kFieldHigh = (1u << (kPosJ - 1)) | (1u << (kPosK - 1)) | (1u << 31);
  constexpr uint32_t kFull = 0xFFFFFFFFu;
  // warp-uniform constants of the coarse lookup (kept opaque: otherwise they are re-derived from `f` inside the loop)
  uint32_t pend = f.pend, cs_i = f.cshift, cs_j = kPosJ + f.cshift, cmask = f.cmask, csI = f.csI, csJ = f.csJ;
  asm volatile("" : "+r"(pend), "+r"(cs_i), "+r"(cs_j), "+r"(cmask), "+r"(csI), "+r"(csJ));
  const uint32_t* __restrict__ coarse = sm.coarse;

  // One lane = one ray.  A lane without a ray is FROZEN (all increments zero, parked on its last cell) and masked out of
  // every side effect by `act`, so that the step below needs no branch for it.
  uint32_t act = 0u, fin = 0u, cbit = 0u;  // fin / cbit, of the step just made: the ray ended without a hit | its new cell may be occupied
  bool exhausted = false, in_face = false, refill = true;
  uint32_t rc = 0, tag = 0, len = 0, L = 0, P = 0, OUT = 0;
  int32_t dPx = 0, dPy = 0, dPz = 0, dLx = 0, dLy = 0, dLz = 0;
  float tx = 0.f, ty = 0.f, tz = 0.f, dx = 0.f, dy = 0.f, dz = 0.f;

  // One step of VoxelGrid::occludedBy for every lane: which axis (VoxelGrid.cpp:299-301,
  // {2,1,2,1,2,2,0,0}[(tx<ty)<<2 | (tx<tz)<<1 | (ty<tz)]: axis 0 iff tx<ty && tx<tz, axis 1 iff !(tx<ty) && ty<tz, else 2), the step
  // itself (:304-305), the memo mark of the cell that is left, the two ways of ending without a hit (:307-308) and the
  // coarse occupancy bit of the cell that is entered.  This is the unit of work of the whole path, and a warp's lanes take
  // different axes all the time: it is written as predicated PTX (the compiler turns the C++ form into branches, which
  // serialise the three axes), ~35 instructions without a branch.  FACE: rays that are still inside the wave's face tell
  // the cells they cross (visited at or after the caster) that this ray is their latest.
  auto step = [&](auto face) {
    if (decltype(face)::value && in_face) {
      const bool lt01 = tx < ty, lt02 = tx < tz, lt12 = ty < tz;
      const bool a0 = lt01 && lt02;
      const bool a1 = !lt01 && lt12;
      const int32_t x = live_rank(f, w, sm, (int32_t)(P & kPosMaskIJ), (int32_t)((P >> kPosJ) & kPosMaskIJ), (int32_t)(P >> kPosK));
      VX_ASSERT(x < C::kMaxLive);
      if (x >= (int32_t)rc) atomicMax(&sm.owner[x], kSet | (tag & kIdMask));
      in_face = !(normal == 0 ? a0 : a1);
    }
    VX_ASSERT(!act || L < f.nvox);
    uint32_t* const mark = f.memo + L;
    asm volatile(
        "{\n\t"
        ".reg .pred a0, a1, a01, on;\n\t"
        "setp.lt.f32 a0, %0, %1;\n\t"
        "setp.lt.and.f32 a0, %0, %2, a0;\n\t"
        "setp.lt.f32 a1, %1, %2;\n\t"
        "setp.geu.and.f32 a1, %0, %1, a1;\n\t"
        "or.pred a01, a0, a1;\n\t"
        "setp.ne.u32 on, %16, 0;\n\t"
        "@on red.relaxed.gpu.global.max.u32 [%14], %15;\n\t"
        "@a0 add.rn.f32 %0, %0, %5;\n\t"
        "@a1 add.rn.f32 %1, %1, %6;\n\t"
        "@!a01 add.rn.f32 %2, %2, %7;\n\t"
        "@a0 add.u32 %3, %3, %8;\n\t"
        "@a1 add.u32 %3, %3, %9;\n\t"
        "@!a01 add.u32 %3, %3, %10;\n\t"
        "@a0 add.u32 %4, %4, %11;\n\t"
        "@a1 add.u32 %4, %4, %12;\n\t"
        "@!a01 add.u32 %4, %4, %13;\n\t"
        "}"
        : "+f"(tx), "+f"(ty), "+f"(tz), "+r"(P), "+r"(L)
        : "f"(dx), "f"(dy), "f"(dz), "r"(dPx), "r"(dPy), "r"(dPz), "r"(dLx), "r"(dLy), "r"(dLz), "l"(mark), "r"(tag), "r"(act)
        : "memory");
    len += act;
    // leaves the grid (pos == Out after a positive step; 0 or wrapped after a negative one: Out = 0 there, and a step to -1
    // ends at the top of the reference loop), or reaches the end voxel: the ray misses
    const uint32_t diff = P ^ OUT;
    fin = ((((diff - kFieldLow) & ~diff & kFieldHigh) != 0u) || (P == pend)) ? 1u : 0u;
    // (a step that left the grid has a wrapped or oversized field in P: the index is masked into the bitmap and the bit it reads
    // is ignored, `fin` decides)
    const uint32_t cw = ((((P >> cs_i) & cmask) << csI) | (((P >> cs_j) & cmask) << csJ) | (P >> (kPosK + 5))) & (C::kCoarseWords - 1u);
    cbit = (coarse[cw] >> ((P >> kPosK) & 31u)) & 1u;
  };

  for (;;) {
    if (refill) {  // warp-uniform: some lane ended its ray (or this is the first round)
      const uint32_t idle = __ballot_sync(kFull, act == 0u);
      if (!exhausted && (idle == kFull || (uint32_t)__popc(idle) >= kRefillLanes)) {
        // the first grab of a warp is its share of the wave's casters (all tracing warps start with rays), later ones fill every idle lane
        const uint32_t take = min((uint32_t)__popc(idle), first_grab);
        first_grab = 32u;
        uint32_t base = 0;
        if (lane == (uint32_t)(__ffs(idle) - 1)) base = atomicAdd(&sm.queue, take);
        base = __shfl_sync(kFull, base, __ffs(idle) - 1);
        if (base + take >= n_cast) exhausted = true;
        const uint32_t mine = (uint32_t)__popc(idle & ((1u << lane) - 1u));
        if (act == 0u && mine < take) {
          const uint32_t s = base + mine;
          if (s < n_cast) {
            VX_ASSERT(s < (uint32_t)C::kMaxLive);
            rc = sm.cand[s];
            VX_ASSERT(rc < (uint32_t)C::kMaxLive);
            tag = enc_epoch | (id_base + s);
            int32_t i, j, k;
            wave_cell(f, w, sm.rcell[rc], i, j, k);
            VxRay r0;
            ray_begin(f, sm, r0, i, j, k);
            tx = r0.tx; dx = r0.dx;
            ty = r0.ty; dy = r0.dy;
            tz = r0.tz; dz = r0.dz;
            dPx = r0.sx;
            dPy = r0.sy * (1 << kPosJ);
            dPz = r0.sz * (1 << kPosK);
            dLx = r0.sx * stride_i;
            dLy = r0.sy * stride_j;
            dLz = r0.sz;
            P = (uint32_t)i | ((uint32_t)j << kPosJ) | ((uint32_t)k << kPosK);
            OUT = (r0.sx > 0 ? (uint32_t)f.g.nx : (i == 0 ? kPosMaskIJ : 0u)) | ((r0.sy > 0 ? (uint32_t)f.g.ny : (j == 0 ? kPosMaskIJ : 0u)) << kPosJ) |
                  ((r0.sz > 0 ? (uint32_t)f.g.nz : (k == 0 ? 0x3FFu : 0u)) << kPosK);
            L = lin(f.g, i, j, k);
            act = 1u;
            in_face = true;
            len = 0;
            ++st.rays;
          }
        }
      }
      if (!__any_sync(kFull, act != 0u)) {
        if (exhausted) break;
        continue;
      }
    }
    // ---- step until some lane needs attention (its ray ended, or the exact occupancy bit has to decide)
    uint32_t attn;
    if (__any_sync(kFull, in_face)) {
      do {
        step(std::true_type{});
        attn = act & (fin | cbit);
      } while (!__any_sync(kFull, attn != 0u) && __any_sync(kFull, in_face));
    } else {
      do {
        step(std::false_type{});
        attn = act & (fin | cbit);
#ifdef VX_TRACE_STATS
        ++st.iters;
#endif
      } while (!__any_sync(kFull, attn != 0u));
    }
#ifdef VX_TRACE_STATS
    ++st.attns;
#endif
    // ---- attention
    bool done = false;
    if (attn) {
      bool hit = false;
      VX_ASSERT(fin || L < f.nvox);
      if (!fin) hit = (__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u;  // ~5 % of the steps: the exact bit decides
      if (fin || hit) {
        VX_ASSERT(((tag & kIdMask) >> 5) < f.hit_words);
        if (hit) {
          vx_red_or(&f.hitbits[(tag & kIdMask) >> 5], 1u << (tag & 31u));
          if (sm.hitcell) sm.hitcell[tag & kIdMask] = L;
        }
#ifdef VX_TRACE_STATS
        atomicMax(&sm.maxlen, len);
#endif
        st.steps += len;
        done = true;
        act = 0u;  // freeze the lane
        in_face = false;
        dx = dy = dz = 0.f;
        dPx = dPy = dPz = 0;
        dLx = dLy = dLz = 0;
      }
    }
    refill = __any_sync(kFull, done);
  }
}

// Runs the columns [a, a + range) of one face as far as one wave reaches; returns the columns consumed.
template <class C>
__device__ int32_t run_wave(const Frame& f, int32_t face, int32_t fixed, int32_t a, int32_t range, VisShared<C>& sm, uint32_t& id_base,
                            Stats& st) {
  const uint32_t tid = threadIdx.x;
  long long t0 = clock64();
  Wave w;
  w.face = face;
  w.fixed = fixed;
  w.a_begin = a;

  // ---- A1: live masks of the chunks of the range, ranks by prefix sum
  const uint32_t n_chunks = (uint32_t)range * (uint32_t)f.cpc;
  uint32_t masks[C::kChunksPerThread];
  uint32_t local = 0;
#pragma unroll
  for (int u = 0; u < C::kChunksPerThread; ++u) {
    const uint32_t c = tid * C::kChunksPerThread + (uint32_t)u;
    masks[u] = 0;
    if (c < n_chunks) {
      uint32_t col, part;
      if (f.cpc == 1) {
        col = c;
        part = 0;
      } else {
        col = c / (uint32_t)f.cpc;
        part = c - col * (uint32_t)f.cpc;
      }
      const int32_t outer = a + (int32_t)col;
      const int32_t i = face == 0 ? fixed : outer, j = face == 0 ? outer : fixed;
      const uint32_t k0 = part * 32u;
      const uint32_t len = min(32u, (uint32_t)f.g.nz - k0);
      masks[u] = extract_bits(f.live, lin(f.g, i, j, (int32_t)k0), len);
    }
    local += (uint32_t)__popc(masks[u]);
  }
  if (tid == 0) {
    sm.ncols_ok = 0;
    sm.any_edge = 0;
    sm.queue = 0;
    sm.maxlen = 0;
  }
  uint32_t base = block_scan(sm, local);
  {
    uint32_t run = base;
#pragma unroll
    for (int u = 0; u < C::kChunksPerThread; ++u) {
      const uint32_t c = tid * C::kChunksPerThread + (uint32_t)u;
      if (c < n_chunks) {
        VX_ASSERT(c < (uint32_t)C::kMaxChunks);
        sm.chmask[c] = masks[u];
        sm.chbase[c] = (uint16_t)(run < 0xFFFFu ? run : 0xFFFFu);
        run += (uint32_t)__popc(masks[u]);
        // last chunk of a column: the column fits if the live cells up to and including it do
        if ((c + 1u) % (uint32_t)f.cpc == 0u && run <= (uint32_t)C::kMaxLive) atomicMax(&sm.ncols_ok, c / (uint32_t)f.cpc + 1u);
      }
    }
  }
  __syncthreads();
  const uint32_t total_live = sm.scan_total;
  if (total_live == 0) return range;  // nothing alive in these columns
  w.ncols = (int32_t)sm.ncols_ok;     // >= 1: one column holds at most nz <= C::kMaxLive live cells
  const uint32_t wave_chunks = (uint32_t)w.ncols * (uint32_t)f.cpc;
  const uint32_t last = wave_chunks - 1u;
  const uint32_t n_live = (uint32_t)sm.chbase[last] + (uint32_t)__popc(sm.chmask[last]);
  if (n_live == 0) return w.ncols;
  st.waves += 1;

  // ---- A2: rank -> cell
#pragma unroll
  for (int u = 0; u < C::kChunksPerThread; ++u) {
    const uint32_t c = tid * C::kChunksPerThread + (uint32_t)u;
    if (c < wave_chunks) {
      uint32_t m = masks[u];
      uint32_t r = sm.chbase[c];
      while (m) {
        const uint32_t b = (uint32_t)__ffs(m) - 1u;
        m &= m - 1u;
        VX_ASSERT(r < (uint32_t)C::kMaxLive);
        sm.rcell[r++] = (uint16_t)((c << 5) | b);
      }
    }
  }
  __syncthreads();

  // ---- A3: snapshot of the memo; candidates in visit order
  const uint32_t per = (n_live + C::kThreads - 1) / C::kThreads;
  const uint32_t r_begin = min(n_live, tid * per), r_end = min(n_live, r_begin + per);
  uint32_t m[C::kLivePerThread];
#pragma unroll
  for (int u = 0; u < C::kLivePerThread; ++u) {
    const uint32_t r = r_begin + (uint32_t)u;
    m[u] = 0;
    if (r < r_end) {
      int32_t i, j, k;
      wave_cell(f, w, sm.rcell[r], i, j, k);
      VX_ASSERT(lin(f.g, i, j, k) < f.nvox);
      m[u] = __ldcg(f.memo + lin(f.g, i, j, k));
    }
  }
  uint32_t cand_flags = 0;
#pragma unroll
  for (int u = 0; u < C::kLivePerThread; ++u) {
    const uint32_t r = r_begin + (uint32_t)u;
    if (r < r_end) {
      if ((m[u] >> kEpochShift) == f.epoch) {
        sm.owner[r] = kSet | (m[u] & kIdMask);
      } else {
        sm.owner[r] = 0;
        cand_flags |= 1u << u;
      }
    }
  }
  uint32_t cslot = block_scan(sm, (uint32_t)__popc(cand_flags));
#pragma unroll
  for (int u = 0; u < C::kLivePerThread; ++u)
    if ((cand_flags >> u) & 1u) {
      VX_ASSERT(cslot < (uint32_t)C::kMaxLive);
      sm.cand[cslot++] = (uint16_t)(r_begin + (uint32_t)u);
    }
  if (tid < C::kMaxLive / 32) {
    sm.killed[tid] = 0;
    sm.decided[tid] = 0;
    sm.castbit[tid] = 0;
    sm.has_succ[tid] = 0;
  }
  __syncthreads();
  const uint32_t n_cand = sm.scan_total;
  {
    const long long t1 = clock64();
    st.cyc[0] += (unsigned long long)(t1 - t0);
    t0 = t1;
  }

  uint32_t n_cast = 0;
  if (n_cand > 0) {
    const int32_t normal = face == 0 ? 0 : 1;
    // ---- B0: cnt[x] = candidates visited before x whose in-face ray crosses x.  The crossed candidates of a thread's
    // first candidate (most waves have <= 128) are remembered in a register, so that the sweeps below do not walk
    // the ray again: up to eight live ranks, 16 bits each.
    unsigned long long crossed = 0, crossed_hi = 0;
    uint32_t n_crossed = 0;  // > 8: more than fit, walk again
    for (uint32_t p = tid; p < n_cand; p += C::kThreads) {
      const uint32_t rc = sm.cand[p];
      int32_t i, j, k;
      wave_cell(f, w, sm.rcell[rc], i, j, k);
      if (normal_first(f, sm, i, j, k, normal)) continue;
      bool any = false;
      prefix_walk(f, w, sm, rc, [&](uint32_t x) {
        VX_ASSERT(x < (uint32_t)C::kMaxLive);
        if (!(sm.owner[x] & kSet)) {
          atomicAdd(&sm.owner[x], 1u);
          any = true;
          if (p == tid) {
            if (n_crossed < 4u) crossed |= (unsigned long long)x << (16u * n_crossed);
            else if (n_crossed < 8u) crossed_hi |= (unsigned long long)x << (16u * (n_crossed - 4u));
            ++n_crossed;
          }
        }
      });
      if (any) {
        atomicOr(&sm.has_succ[p >> 5], 1u << (p & 31u));
        sm.any_edge = 1;
      }
    }
    __syncthreads();
    if (sm.any_edge) {
      // ---- B1: sweeps until every candidate is decided
      st.resolve_waves += 1;
      for (;;) {
        st.rounds += 1;
        bool waiting = false;  // one of this thread's candidates is still undecided after the sweep
        for (uint32_t p = tid; p < n_cand; p += C::kThreads) {
          const uint32_t pbit = 1u << (p & 31u);
          if (sm.decided[p >> 5] & pbit) continue;
          const uint32_t rc = sm.cand[p];
          const bool succ = (sm.has_succ[p >> 5] & pbit) != 0;
          const bool remembered = p == tid && n_crossed <= 8u;
          if ((*(volatile uint32_t*)&sm.killed[rc >> 5] >> (rc & 31u)) & 1u) {
            atomicOr(&sm.decided[p >> 5], pbit);
            if (succ) {  // it does not cast: the cells it would have crossed stop waiting for it
              if (remembered) {
                for (uint32_t q = 0; q < n_crossed; ++q)
                  atomicSub(&sm.owner[(uint32_t)((q < 4u ? crossed : crossed_hi) >> (16u * (q & 3u))) & 0xFFFFu], 1u);
              } else {
                prefix_walk(f, w, sm, rc, [&](uint32_t x) { if (!(sm.owner[x] & kSet)) atomicSub(&sm.owner[x], 1u); });
              }
            }
          } else if (*(volatile uint32_t*)&sm.owner[rc] == 0u) {
            atomicOr(&sm.decided[p >> 5], pbit);
            atomicOr(&sm.castbit[p >> 5], pbit);
            if (succ) {  // it casts: the cells it crosses do not
              if (remembered) {
                for (uint32_t q = 0; q < n_crossed; ++q) {
                  const uint32_t x = (uint32_t)((q < 4u ? crossed : crossed_hi) >> (16u * (q & 3u))) & 0xFFFFu;
                  atomicOr(&sm.killed[x >> 5], 1u << (x & 31u));
                }
              } else {
                prefix_walk(f, w, sm, rc, [&](uint32_t x) { if (!(sm.owner[x] & kSet)) atomicOr(&sm.killed[x >> 5], 1u << (x & 31u)); });
              }
            }
          } else {
            waiting = true;
          }
        }
        if (!__syncthreads_or(waiting ? 1 : 0)) break;  // (one barrier per sweep: it also carries the verdict)
      }
      // cnt[] back to "unset"; casters compacted in visit order (read everything, then write)
      const uint32_t cper = (n_cand + C::kThreads - 1) / C::kThreads;
      const uint32_t p_begin = min(n_cand, tid * cper), p_end = min(n_cand, p_begin + cper);
      uint16_t keep[C::kLivePerThread];
      uint32_t nkeep = 0;
#pragma unroll
      for (int u = 0; u < C::kLivePerThread; ++u) {
        const uint32_t p = p_begin + (uint32_t)u;
        keep[u] = 0;
        if (p < p_end) {
          const uint32_t rc = sm.cand[p];
          sm.owner[rc] = 0;
          if ((sm.castbit[p >> 5] >> (p & 31u)) & 1u) {
            keep[u] = (uint16_t)(rc | 0x8000u);
            ++nkeep;
          }
        }
      }
      uint32_t dst = block_scan(sm, nkeep);  // its barriers separate the reads above from the writes below
#pragma unroll
      for (int u = 0; u < C::kLivePerThread; ++u)
        if (keep[u] & 0x8000u) sm.cand[dst++] = (uint16_t)(keep[u] & 0x7FFFu);
      __syncthreads();
      n_cast = sm.scan_total;
    } else {
      n_cast = n_cand;
    }
    {
      const long long t1 = clock64();
      st.cyc[1] += (unsigned long long)(t1 - t0);
      t0 = t1;
    }

    // ---- C: trace
    uint32_t tracers = (n_cast + C::kRaysPerTracer - 1u) / C::kRaysPerTracer;
    if (tracers > (uint32_t)C::kWarps) tracers = C::kWarps;
    if ((tid >> 5) < tracers) trace_warp(f, w, sm, n_cast, id_base, min(32u, (n_cast + tracers - 1u) / tracers), st);
    __syncthreads();
#ifdef VX_TRACE_STATS
    st.longest += sm.maxlen;
#endif
    {
      const long long t1 = clock64();
      st.cyc[2] += (unsigned long long)(t1 - t0);
      t0 = t1;
    }
  }

  // ---- D: apply the visit.  All hit-bit gathers of a thread are in flight together (they were one
  // dependent L2 / DRAM round trip per cell before).
  {
    uint32_t own[C::kLivePerThread], hw[C::kLivePerThread];
#pragma unroll
    for (int u = 0; u < C::kLivePerThread; ++u) {
      const uint32_t r = tid + (uint32_t)u * C::kThreads;
      own[u] = 0;
      hw[u] = 0;
      if (r < n_live) {
        own[u] = sm.owner[r];
        VX_ASSERT(!(own[u] & kSet) || ((own[u] & kIdMask) >> 5) < f.hit_words);
        if (own[u] & kSet) hw[u] = __ldcg(f.hitbits + ((own[u] & kIdMask) >> 5));
      }
    }
#pragma unroll
    for (int u = 0; u < C::kLivePerThread; ++u) {
      const uint32_t r = tid + (uint32_t)u * C::kThreads;
      if (r >= n_live) break;
      const bool hit = (own[u] & kSet) && ((hw[u] >> (own[u] & 31u)) & 1u);
      if (f.inv && hit) continue;  // the common case: still occluded from this endpoint, nothing to do
      int32_t i, j, k;
      wave_cell(f, w, sm.rcell[r], i, j, k);
      const uint32_t L = lin(f.g, i, j, k);
      if (f.inv) {
        vx_red_and(&f.live[L >> 5], ~(1u << (L & 31u)));  // invalid_ = -1, for good
      } else {
        if (hit) {
          vx_red_or(&f.occl[L >> 5], 1u << (L & 31u));  // last visit wins
          if (sm.hitcell) {  // insertOcclusionLabels needs occlusions_ itself (the occluder's index), not just "> -1"
            const uint32_t Lh = __ldcg(sm.hitcell + (own[u] & kIdMask));
            const uint32_t hk = Lh % (uint32_t)f.g.nz, hij = Lh / (uint32_t)f.g.nz;
            const uint32_t hj = hij % (uint32_t)f.g.ny, hi = hij / (uint32_t)f.g.ny;
            sm.occluder[L] = hi + hj * (uint32_t)f.g.nx + hk * (uint32_t)f.g.nx * (uint32_t)f.g.ny;
          }
        } else {
          vx_red_and(&f.occl[L >> 5], ~(1u << (L & 31u)));
        }
      }
    }
  }
  __syncthreads();
  id_base += n_cast;
  st.cyc[3] += (unsigned long long)(clock64() - t0);
  return w.ncols;
}

template <class C>
__global__ void __launch_bounds__(C::kThreads, C::kCtasPerSm)
vx_visibility_kernel(const VxSlot* __restrict__ slots, const VxVisFrame* __restrict__ frames,
                     const float* __restrict__ endpoints, const VxGridGeom g, const uint32_t hit_words,
                     const VxVisScratch* __restrict__ scratch, const uint32_t n_scratch, uint32_t* locks) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VisShared<C>& sm = *reinterpret_cast<VisShared<C>*>(smem_raw);
  const VxSlot& slot = slots[blockIdx.x];
  const VxVisFrame vf = frames[blockIdx.x];
  const uint32_t tid = threadIdx.x;
  const uint32_t nvox = (uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz;
  const uint32_t words = (nvox + 31u) >> 5;

  unsigned long long t_begin = 0;
  if (tid == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
  // claim a scratch buffer: at most as many CTAs are resident as there are buffers, so one is free
  if (tid == 0) {
    uint32_t s = blockIdx.x % n_scratch;
    while (atomicCAS(&locks[s], 0u, 1u) != 0u) s = s + 1u == n_scratch ? 0u : s + 1u;
    __threadfence();
    sm.scratch_id = s;
  }
  __syncthreads();
  const uint32_t scratch_id = sm.scratch_id;

  Frame f;
  f.g = g;
  f.nvox = nvox;
  f.hit_words = hit_words;
  f.cpc = (g.nz + 31) >> 5;
  f.occ = slot.occ;
  f.live = slot.alive;
  f.memo = scratch[scratch_id].memo;
  f.hitbits = scratch[scratch_id].hitbits;
  f.occl = slot.occl;
  // VX_FRAME_INSERT_OCCLUSION_LABELS (re-read where it is needed: a rare path must not cost the common one a register)
  auto fill_labels_on = [&]() { return (frames[blockIdx.x].flags & 1u) != 0u && slot.occ_filled != nullptr; };
  // coarse bitmap geometry: the smallest block size whose power-of-two padded bitmap fits
  const uint32_t kw = (uint32_t)((g.nz + 31) >> 5);
  uint32_t lkw = 0;
  while ((1u << lkw) < kw) ++lkw;
  f.cshift = kPosJ;  // fallback: a single block
  f.cmask = 0u;
  f.csI = lkw;
  f.csJ = lkw;
  bool coarse_ok = false;
  for (uint32_t s = 0; s < 8u; ++s) {
    const uint32_t cxb = (uint32_t)((g.nx + (1 << s) - 1) >> s), cyb = (uint32_t)((g.ny + (1 << s) - 1) >> s);
    uint32_t lcy = 0;
    while ((1u << lcy) < cyb) ++lcy;
    if (((uint64_t)cxb << (lcy + lkw)) <= (uint64_t)C::kCoarseWords) {
      f.cshift = s;
      f.cmask = kPosMaskIJ >> s;
      f.csJ = lkw;
      f.csI = lcy + lkw;
      coarse_ok = true;
      break;
    }
  }
  // coarse occupancy; live = free cells; occluded starts as "occupied"; no ray has a hit bit yet
  for (uint32_t x = tid; x < C::kCoarseWords; x += C::kThreads) sm.coarse[x] = coarse_ok ? 0u : 0xFFFFFFFFu;
  __syncthreads();
  for (uint32_t x = tid; x < words; x += C::kThreads) {
    const uint32_t o = __ldg(f.occ + x);
    const uint32_t valid = (x == words - 1u && (nvox & 31u)) ? ((1u << (nvox & 31u)) - 1u) : 0xFFFFFFFFu;
    f.live[x] = ~o & valid;
    f.occl[x] = o;
    if (o && coarse_ok) {
      if (g.nz == 32) {  // one word = one column
        const uint32_t i = x / (uint32_t)g.ny, j = x - i * (uint32_t)g.ny;
        atomicOr(&sm.coarse[coarse_word(f, i, j, 0u)], o);
      } else {
        uint32_t bits = o;
        while (bits) {
          const uint32_t b = (uint32_t)__ffs(bits) - 1u;
          bits &= bits - 1u;
          const uint32_t L = x * 32u + b;
          const uint32_t k = L % (uint32_t)g.nz, ij = L / (uint32_t)g.nz;
          const uint32_t i = ij / (uint32_t)g.ny, j = ij - i * (uint32_t)g.ny;
          atomicOr(&sm.coarse[coarse_word(f, i, j, k)], 1u << (k & 31u));
        }
      }
    }
  }
  for (uint32_t x = tid; x < hit_words; x += C::kThreads) f.hitbits[x] = 0;

  Stats st;
  const int32_t shells = vx_num_shells(g.nx, g.ny);
  const int32_t max_cols = C::kMaxChunks / f.cpc;

  const uint32_t n_pass = 1u + vf.n_endpoints;
  for (uint32_t pass = 0; pass < n_pass; ++pass) {
    f.inv = pass > 0;
    if (pass == 0) {
      f.e0 = f.e1 = f.e2 = 0.0f;  // updateOcclusions casts towards the sensor origin
    } else {
      const float* e = endpoints + 3u * (vf.endpoint_offset + pass - 1u);
      f.e0 = e[0];
      f.e1 = e[1];
      f.e2 = e[2];
    }
    f.ex = vx_trunc_i32(VX_FDIV(VX_FSUB(f.e0, g.off[0]), g.res));
    f.ey = vx_trunc_i32(VX_FDIV(VX_FSUB(f.e1, g.off[1]), g.res));
    f.ez = vx_trunc_i32(VX_FDIV(VX_FSUB(f.e2, g.off[2]), g.res));
    f.pend = ((uint32_t)f.ex <= kPosMaskIJ && (uint32_t)f.ey <= kPosMaskIJ && (uint32_t)f.ez <= 0x3FFu)
                 ? ((uint32_t)f.ex | ((uint32_t)f.ey << 11) | ((uint32_t)f.ez << 22))
                 : 0xFFFFFFFFu;
    f.epoch = (pass % kMaxEpoch) + 1u;
    if (tid == 0) {
      const bool record = pass == 0 && fill_labels_on();
      sm.hitcell = record ? scratch[scratch_id].hitcell : nullptr;
      sm.occluder = record ? slot.occluder : nullptr;
    }
    build_axis_table(f, sm);  // (barrier inside)
    if (pass % kMaxEpoch == 0) {  // first pass, or the epoch counter restarts: forget every memo
      uint4* const memo4 = reinterpret_cast<uint4*>(f.memo);  // (the scratch arenas are 256-byte aligned)
      for (uint32_t x = tid; x < (nvox >> 2); x += C::kThreads) memo4[x] = make_uint4(0u, 0u, 0u, 0u);
      for (uint32_t x = (nvox & ~3u) + tid; x < nvox; x += C::kThreads) f.memo[x] = 0u;
    }
    __syncthreads();

    uint32_t id_base = 0;
    for (int32_t o = 0; o < shells; ++o) {
      for (int32_t face = 0; face < 3; ++face) {
        const int32_t outer = face == 0 ? g.ny : g.nx - o - 1;
        const int32_t fixed = face == 0 ? g.nx - 1 - o : (face == 1 ? o : g.ny - 1 - o);
        for (int32_t a = 0; a < outer;) {
          const int32_t range = outer - a < max_cols ? outer - a : max_cols;
          a += run_wave(f, face, fixed, a, range, sm, id_base, st);
        }
      }
    }
    // the ray ids start again at 0: clear the hit bits this pass used
    const uint32_t used_words = min(hit_words, (id_base + 31u) >> 5);
    for (uint32_t x = tid; x < used_words; x += C::kThreads) f.hitbits[x] = 0;
    if (pass == 0) {
      const bool fill_labels = fill_labels_on();
      if (fill_labels) {
        // VoxelGrid::insertOcclusionLabels (VoxelGrid.cpp:75-96): a voxel that was empty when the occlusions were computed takes
        // count and labels of the first occupied voxel above it if only occluded empty voxels lie between.  The reference walks
        // each column bottom-up and looks upwards at counts it has not modified yet: one top-down scan per column with the
        // ORIGINAL occupancy is the same thing.  Filled voxels are occupied for every pass that follows (occ_filled); occluded
        // ones keep invalid_ = min(occluder index, own index) (VoxelGrid.cpp:187-190,227): isInvalid iff the occluder's
        // reference-internal index is the smaller one -- or, without any updateInvalid pass, iff occluded at all.
        __syncthreads();  // occl is final
        for (uint32_t x = tid; x < words; x += C::kThreads) {
          slot.occ_filled[x] = __ldg(f.occ + x);
          slot.out_invalid[x] = 0u;  // (scratch until the pack kernel: filled voxels that stay invalid)
        }
        __syncthreads();
        const uint32_t none = 0xFFFFFFFFu;
        for (uint32_t col = tid; col < (uint32_t)g.nx * (uint32_t)g.ny; col += C::kThreads) {
          const int32_t ci = (int32_t)(col / (uint32_t)g.ny), cj = (int32_t)(col % (uint32_t)g.ny);
          uint32_t src = none;
          for (int32_t ck = g.nz - 1; ck >= 0; --ck) {
            const uint32_t L = lin(g, ci, cj, ck);
            const bool occupied_here = (__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u;
            const bool occluded_here = (__ldcg(f.occl + (L >> 5)) >> (L & 31u)) & 1u;
            uint32_t from = none;
            if (occupied_here) {
              src = L;
            } else {
              from = src;
              if (!occluded_here) src = none;  // a visible empty voxel ends the run for everything below it
            }
            slot.fill_source[L] = from;
            if (from != none) {
              slot.out_label[L] = slot.out_label[from];
              atomicOr(&slot.occ_filled[L >> 5], 1u << (L & 31u));
              const uint32_t own_index = (uint32_t)ci + (uint32_t)cj * (uint32_t)g.nx + (uint32_t)ck * (uint32_t)g.nx * (uint32_t)g.ny;
              if (occluded_here && (n_pass == 1u || slot.occluder[L] < own_index)) atomicOr(&slot.out_invalid[L >> 5], 1u << (L & 31u));
            }
          }
        }
        __syncthreads();
        f.occ = slot.occ_filled;  // first read from here on (never cached before): what the invalid passes see
        for (uint32_t x = tid; x < words; x += C::kThreads) {  // the filled voxels join the coarse bitmap
          const uint32_t added = __ldcg(slot.occ_filled + x) & ~__ldg(slot.occ + x);
          uint32_t bits = added;
          while (bits) {
            const uint32_t b = (uint32_t)__ffs(bits) - 1u;
            bits &= bits - 1u;
            const uint32_t L = x * 32u + b;
            const uint32_t k = L % (uint32_t)g.nz, ij = L / (uint32_t)g.nz;
            const uint32_t i = ij / (uint32_t)g.ny, j = ij - i * (uint32_t)g.ny;
            atomicOr(&sm.coarse[coarse_word(f, i, j, k)], 1u << (k & 31u));
          }
        }
      }
      // invalid_ = occlusions_ (VoxelGrid.cpp:169): what is not occluded is dead from here on
      uint32_t n_live0 = 0, n_occ = 0;
      for (uint32_t x = tid; x < words; x += C::kThreads) {
        const uint32_t o = fill_labels ? __ldcg(slot.occ_filled + x) : __ldg(f.occ + x);
        const uint32_t l = __ldcg(f.occl + x) & ~o;
        f.live[x] = l;
        n_live0 += (uint32_t)__popc(l);
        n_occ += (uint32_t)__popc(o);
      }
      atomicAdd(&slot.counters[12], (unsigned long long)n_live0 | ((unsigned long long)n_occ << 32));  // (statistics)
    }
    __syncthreads();
  }

  if (fill_labels_on()) {  // filled voxels that stay invalid; `.bin` of a grid that is its own input grid
    for (uint32_t x = tid; x < words; x += C::kThreads) f.live[x] = __ldcg(f.live + x) | slot.out_invalid[x];
    if (frames[blockIdx.x].flags & 2u) {  // VX_FRAME_INPUT_IS_TARGET
      for (uint32_t x = tid; x < words; x += C::kThreads) {
        uint32_t bits = 0u;
        for (uint32_t b = 0; b < 32u; ++b) {
          const uint32_t L = x * 32u + b;
          if (L < nvox && slot.out_label[L] > 0) bits |= 1u << b;
        }
        slot.out_bin[x] = __byte_perm(__brev(bits), 0u, 0x0123u);  // MSB-first bytes (voxelize_utils.cpp:205-216)
      }
    }
    __syncthreads();
  }

  // statistics: rays / steps summed over threads, the rest as seen by thread 0
  unsigned long long rays = st.rays, steps = st.steps;
  for (int off = 16; off > 0; off >>= 1) {
    rays += __shfl_down_sync(0xFFFFFFFFu, rays, off);
    steps += __shfl_down_sync(0xFFFFFFFFu, steps, off);
  }
  if ((tid & 31u) == 0) {
    atomicAdd(&slot.counters[0], rays);
    atomicAdd(&slot.counters[1], steps);
  }
  __syncthreads();
  if (tid == 0) {
    __threadfence();
    atomicExch(&locks[scratch_id], 0u);  // release the scratch buffer
    slot.counters[2] = st.waves;
    slot.counters[3] = st.resolve_waves;
    slot.counters[4] = st.rounds;
    for (int k = 0; k < 4; ++k) slot.counters[5 + k] = st.cyc[k];
    slot.counters[13] = st.longest;
    slot.counters[14] = st.iters;
    slot.counters[15] = st.attns;
    // where and when the frame ran (developer statistics: VX_B200_FRAME_STATS)
    uint32_t smid;
    unsigned long long t_end;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
    slot.counters[9] = smid;
    slot.counters[10] = t_begin;
    slot.counters[11] = t_end;
  }
}

// VoxelGrid::occludedBy(i, j, k, endpoint, visited*) as a PUBLIC call (VoxelGrid.h:89-90, VoxelGrid.cpp:235-316): one ray,
// every cell the reference loop looks at (the occupied one that ends it included), and the reference-internal index
// i + j * Sx + k * Sx * Sy of the occluder or -1.  The memo side effect of the reference is not observable through the
// public API (both passes reset occludedBy_ before they read it), so there is none here.
__global__ void vx_trace_ray_kernel(const uint32_t* __restrict__ occ, const VxGridGeom g, const int32_t i, const int32_t j,
                                    const int32_t k, const float e0, const float e1, const float e2, int32_t* visited,
                                    const uint32_t capacity, int32_t* result) {
  VxRay r;
  r.begin(g, i, j, k, e0, e1, e2);
  uint32_t n = 0;
  int32_t occluder = -1;
  for (;;) {
    if (!r.inside(g)) break;
    if (n < capacity) {
      visited[3u * n + 0u] = r.i;
      visited[3u * n + 1u] = r.j;
      visited[3u * n + 2u] = r.k;
    }
    ++n;
    const uint32_t L = lin(g, r.i, r.j, r.k);
    if ((occ[L >> 5] >> (L & 31u)) & 1u) {
      occluder = r.i + r.j * g.nx + r.k * g.nx * g.ny;
      break;
    }
    if (!r.advance()) break;
  }
  result[0] = (int32_t)n;
  result[1] = occluder;
}

}  // namespace

cudaError_t vx_launch_trace_ray(const uint32_t* occ, VxGridGeom geom, int32_t i, int32_t j, int32_t k, const float* endpoint,
                                int32_t* visited, uint32_t capacity, int32_t* result, cudaStream_t stream) {
  vx_trace_ray_kernel<<<1, 1, 0, stream>>>(occ, geom, i, j, k, endpoint[0], endpoint[1], endpoint[2], visited, capacity, result);
  return cudaGetLastError();
}

// 7 x (30 KB + 1 KB) of shared memory and 7 x 128 x 72 registers per SM | one CTA of 512 threads (92 KB) per SM
#ifndef VX_THR_THREADS  // (developer switches for A/B builds of the throughput shape: tools/build_variants.py)
#define VX_THR_THREADS 128
#define VX_THR_MAXLIVE 1280
#define VX_THR_CTAS 7
#define VX_THR_COARSE 4096
#endif
using CfgThroughput = VisCfg<VX_THR_THREADS, VX_THR_MAXLIVE, 256, VX_THR_CTAS, VX_THR_COARSE>;
#ifndef VX_LAT_THREADS
#define VX_LAT_THREADS 512
#endif
using CfgLatency = VisCfg<VX_LAT_THREADS, 8192, 1024, 1>;
// MIDDLE shape: 256 threads, waves of <= 2560 live cells, 4 CTAs per SM (64 registers) -- for batches that give every SM two to
// four frames (a rank's share of a sequence split over eight GPUs: 568 frames): twice the threads per frame of the throughput
// shape on the same register file.
#ifndef VX_MID_THREADS
#define VX_MID_THREADS 256
#define VX_MID_MAXLIVE 2560
#define VX_MID_CTAS 4
#endif
using CfgMid = VisCfg<VX_MID_THREADS, VX_MID_MAXLIVE, 512, VX_MID_CTAS>;

constexpr size_t kRoomySmem = 36 * 1024;  // 6 x (36 + 1) KB fit an SM, 7 do not

cudaError_t vx_visibility_prepare() {
  cudaError_t e = cudaFuncSetAttribute(vx_visibility_kernel<CfgThroughput>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)std::max(sizeof(VisShared<CfgThroughput>), kRoomySmem));
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(vx_visibility_kernel<CfgMid>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VisShared<CfgMid>));
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(vx_visibility_kernel<CfgLatency>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VisShared<CfgLatency>));
}

// CTAs of the kernel the device holds at once (the larger of the two shapes: the scratch pool is sized by it)
int vx_visibility_resident_ctas(int device) {
  int a = 0, b = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&a, vx_visibility_kernel<CfgThroughput>, CfgThroughput::kThreads,
                                                    sizeof(VisShared<CfgThroughput>)) != cudaSuccess) return 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&b, vx_visibility_kernel<CfgLatency>, CfgLatency::kThreads,
                                                    sizeof(VisShared<CfgLatency>)) != cudaSuccess) return 0;
  int m = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&m, vx_visibility_kernel<CfgMid>, CfgMid::kThreads, sizeof(VisShared<CfgMid>)) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return std::max(a, std::max(b, m)) * sms;
}

// mode: 0 = by batch size (latency shape when the frames cannot even give every SM two CTAs, middle shape up to four per SM and
// for batches a little larger than the throughput shape's 7 per SM), 1 = throughput, 2 = latency, 3 = middle.
// leave_room: another batch is in flight or queued -- the throughput shape then asks for a little more shared memory than it uses,
// so that an SM holds 6 instead of 7 of its CTAs and keeps registers for one CTA of the other batch's fill / vote / emit kernels
// (which otherwise crawl behind ~900 resident traversal CTAs: 294 / 175 / 478 ms instead of 40 / 3 / 6 per step).
cudaError_t vx_launch_visibility(const VxSlot* slots, const VxVisFrame* frames, uint32_t n_slots,
                                 const float* endpoints, VxGridGeom geom, const VxVisScratch* scratch, uint32_t n_scratch,
                                 uint32_t* locks, int mode, int leave_room, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  const uint32_t hit_words = (uint32_t)(vx_visibility_visits(geom) / 32u + 1u);
  if (mode == 0) {
    int dev = 0, sms = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    static const bool mid_auto = [] { const char* e = std::getenv("VX_B200_MID_SHAPE"); return !e || std::atoi(e) != 0; }();  // (=0: A/B)
    // Measured (profiles/r02_middle_shape_ab.txt): the middle shape wins where the throughput shape would run with few frames
    // per SM (568 frames: 1.23 s against 1.73 s) and where it would need a second, nearly empty round (1136 frames on 1036
    // slots: 2.40 s against 3.00 s); the throughput shape wins when everything is resident at once (909 frames: 2.02 s against
    // 2.30 s); for many rounds the two are equal (489 / 495 frames/s under streaming load).
    const uint32_t thr_slots = (uint32_t)CfgThroughput::kCtasPerSm * (uint32_t)sms, mid_slots = (uint32_t)CfgMid::kCtasPerSm * (uint32_t)sms;
    const bool mid = mid_auto && (n_slots <= mid_slots || (n_slots > thr_slots && n_slots <= mid_slots * 5u / 2u));
    mode = n_slots <= 2u * (uint32_t)sms ? 2 : (mid ? 3 : 1);
  }
  if (mode == 3)
    vx_visibility_kernel<CfgMid><<<n_slots, CfgMid::kThreads, sizeof(VisShared<CfgMid>), stream>>>(
        slots, frames, endpoints, geom, hit_words, scratch, n_scratch, locks);
  else if (mode == 2)
    vx_visibility_kernel<CfgLatency><<<n_slots, CfgLatency::kThreads, sizeof(VisShared<CfgLatency>), stream>>>(
        slots, frames, endpoints, geom, hit_words, scratch, n_scratch, locks);
  else {
    size_t smem = sizeof(VisShared<CfgThroughput>);
    if (leave_room && smem < kRoomySmem) smem = kRoomySmem;
    vx_visibility_kernel<CfgThroughput><<<n_slots, CfgThroughput::kThreads, smem, stream>>>(
        slots, frames, endpoints, geom, hit_words, scratch, n_scratch, locks);
  }
  return cudaGetLastError();
}

uint64_t vx_visibility_visits(VxGridGeom g) {
  uint64_t v = 0;
  const int32_t shells = vx_num_shells(g.nx, g.ny);
  for (int32_t o = 0; o < shells; ++o) v += (uint64_t)g.nz * ((uint64_t)g.ny + 2ull * (uint64_t)(g.nx - o - 1));
  return v;
}

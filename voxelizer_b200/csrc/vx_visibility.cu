// K4: the visibility passes -- VoxelGrid::updateOcclusions (VoxelGrid.cpp:98-171), one
// VoxelGrid::updateInvalid per past scan (VoxelGrid.cpp:173-233) and the memoised ray they share,
// VoxelGrid::occludedBy (VoxelGrid.cpp:235-316) -- reproduced bit for bit.
//
// The reference is order dependent: a ray writes its result into occludedBy_ for every cell it
// crossed, unconditionally, and a cell visited later reads whatever the LATEST ray that crossed it
// left there (casting its own ray only if nothing has crossed it yet).  Casting one independent ray
// per voxel flips ~60 k output bits per frame.  This kernel keeps the reference's visit order as a
// timestamp and parallelises inside it:
//
//   * visit number tau: position of a cell visit in the reference's shell / face / column / z order.
//   * memo[L] = (epoch << 25) | tau_ray, merged with atomicMax: what a visit reads is the ray with the
//     largest tau that crossed the cell = the latest writer; that ray's result (hit / miss) is looked
//     up in rayhit[tau_ray], written once by the ray when it ends -- so a ray is walked ONCE.
//     `epoch` = pass number, so the reference's "fill occludedBy_ with -2" costs nothing.
//     0xFFFFFFFF = the cell is DEAD (invalid_ == -1): every later pass skips it, atomicMax keeps it dead.
//   * a WAVE = a tau-contiguous run of whole columns of one face (<= kWaveCells cells).  One CTA owns
//     one frame and runs its waves in order; inside a wave
//       A  snapshot the memo of every cell; candidates = visited cells nothing has crossed yet;
//       B  decide which candidates cast: c casts iff no CASTING candidate with smaller tau crosses c
//          with the in-face prefix of its ray (rays are monotone per axis, so they never re-enter a
//          face).  Monotone fix-point over the undecided candidates, all in shared memory;
//       C  casters trace their rays in parallel (atomicMax into memo; cells of the same wave that
//          are visited at or after the caster also merge into the snapshot);
//       D  apply the visit rule from the snapshot: occlusions_[c] = memo (last visit wins) or
//          invalid_[c] = min(invalid_[c], memo)  (a miss kills the cell).
//   * only booleans are observable (isOccluded / isInvalid, VoxelGrid.cpp:65-73): occluded = hit,
//     alive = hit at every visit of every pass.
// tests/model/wave_model.cpp is the CPU model of exactly this formulation.
//
// Latency-bound, not HBM-bound.  The occupancy a ray tests at every step is read-only for the whole
// kernel and sparse (a few % of the voxels), so each CTA first compresses it into shared memory:
// one bit per 4x4x2 block, a rank directory and one 32-bit fine mask per non-empty block; a lookup
// is 1-3 shared-memory loads instead of an L2 round trip.  What stays in L2/HBM is the memo: one
// coalesced read per visited cell and one fire-and-forget RED.MAX per ray step.
#include "vx_launch.h"

namespace {

constexpr int kVisThreads = 256;
constexpr int kWaveCells = 4096;
constexpr int kCellsPerThread = kWaveCells / kVisThreads;
constexpr uint32_t kSkip = 0xFFFFFFFFu;
constexpr uint32_t kDead = 0xFFFFFFFFu;
constexpr uint32_t kEpochShift = 25;
constexpr uint32_t kTauMask = (1u << kEpochShift) - 1u;
constexpr uint32_t kMaxEpoch = 126;  // epochs 1..126: an encoded memo can never equal kDead

// sparse occupancy in shared memory
constexpr uint32_t kCoarseWordsMax = 2048;  // 65536 blocks of 4x4x2 voxels (256 x 256 x 32 grid)
constexpr uint32_t kFineCap = 8192;         // non-empty blocks held in shared memory; the rest falls back to L2

struct VisShared {
  uint32_t snap[kWaveCells];  // 0 unset | kSkip | 2 + hit (crossed in an earlier wave) | (q + 2) << 1 (in-wave caster q)
  uint16_t cand[kWaveCells];  // candidate list (local indices)
  uint8_t status[kWaveCells]; // per candidate slot: 0 undecided, 1 casts, 2 does not
  uint32_t cov_und[kWaveCells / 32];   // cell is crossed by an UNDECIDED earlier candidate
  uint32_t cov_cast[kWaveCells / 32];  // cell is crossed by a DECIDED earlier caster
  uint32_t hitbits[kWaveCells / 32];   // result of in-wave caster q
  uint32_t coarse[kCoarseWordsMax];    // bit per 4x4x2 block: any voxel occupied
  uint16_t rank[kCoarseWordsMax];      // number of set coarse bits before this word (saturating)
  uint32_t fine[kFineCap];             // occupancy of the 32 voxels of the r-th non-empty block
  uint32_t n_cand;
  uint32_t any_forward;
  uint32_t remaining;
  uint32_t scan_carry;
};

struct Wave {
  int32_t face;     // 0: i fixed, outer j; 1/2: j fixed, outer i
  int32_t fixed;    // the fixed coordinate
  int32_t a_begin;  // first outer coordinate
  int32_t ncols;
  uint32_t tau_base;
};

struct Stats {
  unsigned long long rays = 0, steps = 0, waves = 0, resolve_waves = 0, rounds = 0;
  unsigned long long cyc[4] = {0, 0, 0, 0};
};

struct Frame {
  VxGridGeom g;
  int32_t nz_shift;  // log2(nz) if nz is a power of two, else -1
  int32_t cy, cz;    // coarse grid dims (y, z); 0 = sparse occupancy disabled (grid too large)
  const uint32_t* __restrict__ occ;
  uint32_t* memo;
  uint8_t* rayhit;  // [visits per pass] result of the ray cast at visit tau (this pass)
  uint32_t* occl;
  float e0, e1, e2;
  uint32_t epoch;
  bool inv;
};

__device__ __forceinline__ uint32_t lin(const VxGridGeom& g, int32_t i, int32_t j, int32_t k) {
  return ((uint32_t)i * (uint32_t)g.ny + (uint32_t)j) * (uint32_t)g.nz + (uint32_t)k;
}

// occupancy test of an in-grid cell
__device__ __forceinline__ bool occupied(const Frame& f, const VisShared& sm, int32_t i, int32_t j, int32_t k) {
  if (f.cz > 0) {
    const uint32_t C = ((uint32_t)(i >> 2) * (uint32_t)f.cy + (uint32_t)(j >> 2)) * (uint32_t)f.cz + (uint32_t)(k >> 1);
    const uint32_t word = sm.coarse[C >> 5];
    const uint32_t bit = 1u << (C & 31u);
    if (!(word & bit)) return false;
    const uint32_t r = (uint32_t)sm.rank[C >> 5] + (uint32_t)__popc(word & (bit - 1u));
    if (r < kFineCap) return (sm.fine[r] >> ((((uint32_t)i & 3u) * 4u + ((uint32_t)j & 3u)) * 2u + ((uint32_t)k & 1u))) & 1u;
  }
  const uint32_t L = lin(f.g, i, j, k);
  return (__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u;
}

// Builds the sparse occupancy of this frame in shared memory (once per frame).
__device__ void build_sparse_occupancy(const Frame& f, VisShared& sm) {
  const uint32_t tid = threadIdx.x;
  const int32_t cx = (f.g.nx + 3) >> 2;
  const uint32_t n_blocks = (uint32_t)cx * (uint32_t)f.cy * (uint32_t)f.cz;
  const uint32_t n_words = (n_blocks + 31u) >> 5;
  auto block_mask = [&](uint32_t C) -> uint32_t {
    const uint32_t ck = C % (uint32_t)f.cz;
    const uint32_t cj = (C / (uint32_t)f.cz) % (uint32_t)f.cy;
    const uint32_t ci = C / ((uint32_t)f.cz * (uint32_t)f.cy);
    uint32_t m = 0;
    for (uint32_t di = 0; di < 4; ++di)
      for (uint32_t dj = 0; dj < 4; ++dj)
        for (uint32_t dk = 0; dk < 2; ++dk) {
          const int32_t i = (int32_t)(ci * 4 + di), j = (int32_t)(cj * 4 + dj), k = (int32_t)(ck * 2 + dk);
          if (i < f.g.nx && j < f.g.ny && k < f.g.nz) {
            const uint32_t L = lin(f.g, i, j, k);
            m |= ((__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u) << ((di * 4 + dj) * 2 + dk);
          }
        }
    return m;
  };
  // pass 1: coarse bits (one thread per coarse word: 32 consecutive blocks)
  for (uint32_t w = tid; w < n_words; w += kVisThreads) {
    uint32_t bits = 0;
    for (uint32_t b = 0; b < 32; ++b) {
      const uint32_t C = w * 32 + b;
      if (C < n_blocks && block_mask(C)) bits |= 1u << b;
    }
    sm.coarse[w] = bits;
  }
  if (tid == 0) sm.scan_carry = 0;
  __syncthreads();
  // pass 2: exclusive prefix sum of popcounts over the coarse words (chunks of kVisThreads words)
  for (uint32_t w0 = 0; w0 < n_words; w0 += kVisThreads) {
    const uint32_t w = w0 + tid;
    const uint32_t c = w < n_words ? (uint32_t)__popc(sm.coarse[w]) : 0u;
    // block-wide inclusive scan via warp shuffles + the cand[] scratch for warp totals
    uint32_t x = c;
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t y = __shfl_up_sync(0xFFFFFFFFu, x, off);
      if ((tid & 31u) >= (uint32_t)off) x += y;
    }
    uint32_t* warp_tot = sm.cov_und;  // scratch (kVisThreads / 32 entries)
    if ((tid & 31u) == 31u) warp_tot[tid >> 5] = x;
    __syncthreads();
    uint32_t before = sm.scan_carry;
    for (uint32_t ww = 0; ww < (tid >> 5); ++ww) before += warp_tot[ww];
    const uint32_t excl = before + x - c;
    if (w < n_words) sm.rank[w] = (uint16_t)(excl < 0xFFFFu ? excl : 0xFFFFu);
    __syncthreads();
    if (tid == kVisThreads - 1) sm.scan_carry = before + x;
    __syncthreads();
  }
  // pass 3: fine masks of the non-empty blocks that fit
  for (uint32_t w = tid; w < n_words; w += kVisThreads) {
    uint32_t bits = sm.coarse[w];
    uint32_t r = sm.rank[w];
    while (bits) {
      const uint32_t b = (uint32_t)__ffs(bits) - 1u;
      bits &= bits - 1u;
      if (r < kFineCap) sm.fine[r] = block_mask(w * 32 + b);
      ++r;
    }
  }
  __syncthreads();
}

__device__ __forceinline__ void cell_of(const Frame& f, const Wave& w, uint32_t q, int32_t& i, int32_t& j, int32_t& k) {
  uint32_t col;
  if (f.nz_shift >= 0) {
    col = q >> f.nz_shift;
    k = (int32_t)(q & ((1u << f.nz_shift) - 1u));
  } else {
    col = q / (uint32_t)f.g.nz;
    k = (int32_t)(q - col * (uint32_t)f.g.nz);
  }
  if (w.face == 0) {
    i = w.fixed;
    j = w.a_begin + (int32_t)col;
  } else {
    i = w.a_begin + (int32_t)col;
    j = w.fixed;
  }
}

// local index of (i,j,k) inside the wave, or -1
__device__ __forceinline__ int32_t local_of(const Frame& f, const Wave& w, int32_t i, int32_t j, int32_t k) {
  const int32_t fx = w.face == 0 ? i : j;
  const int32_t outer = (w.face == 0 ? j : i) - w.a_begin;
  if (fx != w.fixed || outer < 0 || outer >= w.ncols) return -1;
  return outer * f.g.nz + k;
}

// Cells of this wave, visited after candidate qc, that qc's ray would cross before it leaves the
// face, hits an occupied cell or terminates: sets their bit in `target`.  Only LATER cells are ever
// marked, so "bit set" means "crossed by an earlier candidate".  Returns #cells marked.
__device__ uint32_t prefix_mark(const Frame& f, const Wave& w, const VisShared& sm, uint32_t qc, uint32_t* target) {
  int32_t i, j, k;
  cell_of(f, w, qc, i, j, k);
  if (occupied(f, sm, i, j, k)) return 0;  // an occupied start cell traverses nothing
  VxRay r;
  r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
  const int32_t normal = w.face == 0 ? 0 : 1;
  uint32_t marked = 0;
  for (;;) {
    if (r.next_axis() == normal) break;
    if (!r.advance()) break;
    if (!r.inside(f.g)) break;
    if (occupied(f, sm, r.i, r.j, r.k)) break;
    const int32_t q = local_of(f, w, r.i, r.j, r.k);
    if (q > (int32_t)qc) {
      atomicOr(&target[q >> 5], 1u << (q & 31));
      ++marked;
    }
  }
  return marked;
}

// Phase C for one caster: the reference's occludedBy(), walked once.
__device__ void trace(const Frame& f, const Wave& w, VisShared& sm, uint32_t qc, unsigned long long& steps) {
  int32_t i, j, k;
  cell_of(f, w, qc, i, j, k);
  const uint32_t tau = w.tau_base + qc;
  const uint32_t enc = (f.epoch << kEpochShift) | tau;
  const uint32_t senc = (qc + 2u) << 1;
  bool hit = false;
  if (occupied(f, sm, i, j, k)) {  // returns its own index; the caller's assignment memoises it (VoxelGrid.cpp:112)
    atomicMax(&f.memo[lin(f.g, i, j, k)], enc);
    atomicMax(&sm.snap[qc], senc);
    hit = true;
  } else {
    VxRay r;
    r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
    uint32_t len = 0;
    for (;;) {
      if (!r.inside(f.g)) break;
      if (occupied(f, sm, r.i, r.j, r.k)) {
        hit = true;
        break;
      }
      atomicMax(&f.memo[lin(f.g, r.i, r.j, r.k)], enc);
      const int32_t q = local_of(f, w, r.i, r.j, r.k);
      if (q >= (int32_t)qc) atomicMax(&sm.snap[q], senc);
      ++len;
      if (!r.advance()) break;
    }
    steps += len;
  }
  f.rayhit[tau] = hit ? 1 : 0;
  if (hit) atomicOr(&sm.hitbits[qc >> 5], 1u << (qc & 31u));
}

__device__ void run_wave(const Frame& f, const Wave& w, VisShared& sm, Stats& st) {
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t n = (uint32_t)w.ncols * (uint32_t)f.g.nz;

  if (tid == 0) {
    sm.n_cand = 0;
    sm.any_forward = 0;
    sm.remaining = 0;
  }
  if (tid < kWaveCells / 32) {
    sm.cov_und[tid] = 0;
    sm.cov_cast[tid] = 0;
    sm.hitbits[tid] = 0;
  }
  long long t0 = clock64();
  st.waves += 1;

  // ---- A: snapshot + candidates.  All memo loads of a thread are issued before any is consumed.
  uint32_t m[kCellsPerThread];
#pragma unroll
  for (int u = 0; u < kCellsPerThread; ++u) {
    const uint32_t q = (uint32_t)u * kVisThreads + tid;
    m[u] = 0;
    if (q < n) {
      int32_t i, j, k;
      cell_of(f, w, q, i, j, k);
      m[u] = __ldcg(f.memo + lin(f.g, i, j, k));
    }
  }
  __syncthreads();  // n_cand / bitmaps initialised
  // result of the ray that crossed each already-crossed cell (second, dependent gather)
  uint32_t h[kCellsPerThread];
#pragma unroll
  for (int u = 0; u < kCellsPerThread; ++u) {
    h[u] = 0;
    if (m[u] != kDead && (m[u] >> kEpochShift) == f.epoch) h[u] = 2u | (uint32_t)__ldcg(f.rayhit + (m[u] & kTauMask));
  }
#pragma unroll
  for (int u = 0; u < kCellsPerThread; ++u) {
    const uint32_t q = (uint32_t)u * kVisThreads + tid;
    bool is_cand = false;
    if (q < n) {
      uint32_t s;
      if (f.inv && m[u] == kDead) {
        s = kSkip;  // invalid_ == -1: "just skip" (VoxelGrid.cpp:187)
      } else if (h[u]) {
        s = h[u];
      } else {
        s = 0u;
        is_cand = true;
      }
      sm.snap[q] = s;
    }
    const uint32_t ballot = __ballot_sync(0xFFFFFFFFu, is_cand);
    if (ballot) {
      uint32_t base = 0;
      if (lane == 0) base = atomicAdd(&sm.n_cand, (uint32_t)__popc(ballot));
      base = __shfl_sync(0xFFFFFFFFu, base, 0);
      if (is_cand) sm.cand[base + __popc(ballot & ((1u << lane) - 1u))] = (uint16_t)q;
    }
  }
  __syncthreads();
  const uint32_t n_cand = sm.n_cand;
  {
    const long long t1 = clock64();
    st.cyc[0] += (unsigned long long)(t1 - t0);
    t0 = t1;
  }

  if (n_cand > 0) {
    // ---- B: which candidates cast?
    for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
      sm.status[p] = 0;
      if (prefix_mark(f, w, sm, sm.cand[p], sm.cov_und)) sm.any_forward = 1;
    }
    __syncthreads();
    if (sm.any_forward) {
      st.resolve_waves += 1;
      for (;;) {
        st.rounds += 1;
        // decide what can be decided: crossed by a decided caster -> does not cast; crossed by no
        // undecided candidate -> casts (and marks what it crosses); otherwise wait for a later round
        for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
          if (sm.status[p] != 0) continue;
          const uint32_t qc = sm.cand[p];
          const uint32_t bit = 1u << (qc & 31u);
          if (sm.cov_cast[qc >> 5] & bit) {
            sm.status[p] = 2;
          } else if (!(sm.cov_und[qc >> 5] & bit)) {
            sm.status[p] = 1;
            prefix_mark(f, w, sm, qc, sm.cov_cast);
          } else {
            sm.remaining = 1;
          }
        }
        __syncthreads();
        if (!sm.remaining) break;
        __syncthreads();
        if (tid == 0) sm.remaining = 0;
        if (tid < kWaveCells / 32) sm.cov_und[tid] = 0;
        __syncthreads();
        for (uint32_t p = tid; p < n_cand; p += kVisThreads)
          if (sm.status[p] == 0) prefix_mark(f, w, sm, sm.cand[p], sm.cov_und);
        __syncthreads();
      }
    } else {
      for (uint32_t p = tid; p < n_cand; p += kVisThreads) sm.status[p] = 1;
      __syncthreads();
    }
    {
      const long long t1 = clock64();
      st.cyc[1] += (unsigned long long)(t1 - t0);
      t0 = t1;
    }

    // ---- C: trace
    for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
      if (sm.status[p] != 1) continue;
      ++st.rays;
      trace(f, w, sm, sm.cand[p], st.steps);
    }
    __syncthreads();
    {
      const long long t1 = clock64();
      st.cyc[2] += (unsigned long long)(t1 - t0);
      t0 = t1;
    }
  }

  // ---- D: apply the visit
  for (int u = 0; u < kCellsPerThread; ++u) {
    const uint32_t q = (uint32_t)u * kVisThreads + tid;
    if ((uint32_t)u * kVisThreads + (tid & ~31u) >= n) break;  // warp-uniform
    bool visited = false, hit = false;
    uint32_t L = 0;
    if (q < n) {
      const uint32_t s = sm.snap[q];
      if (s != kSkip) {
        int32_t i, j, k;
        cell_of(f, w, q, i, j, k);
        L = lin(f.g, i, j, k);
        const uint32_t rk = s >> 1;  // 1: crossed before this wave (hit in bit 0); >= 2: in-wave caster rk - 2
        hit = rk >= 2u ? ((sm.hitbits[(rk - 2u) >> 5] >> ((rk - 2u) & 31u)) & 1u) != 0 : (s & 1u) != 0;
        visited = true;
      }
    }
    if (f.inv) {
      if (visited && !hit) atomicMax(&f.memo[L], kDead);  // invalid_ = -1, for good
    } else {
      const uint32_t active = __ballot_sync(0xFFFFFFFFu, visited);
      if (visited) {
        const uint32_t word = L >> 5;
        const uint32_t bit = 1u << (L & 31u);
        const uint32_t peers = __match_any_sync(active, word);
        const uint32_t vis = __reduce_or_sync(peers, bit);
        const uint32_t hits = __reduce_or_sync(peers, hit ? bit : 0u);
        if ((uint32_t)(__ffs(peers) - 1) == lane) {
          if (vis != hits) atomicAnd(&f.occl[word], ~(vis & ~hits));
          if (hits) atomicOr(&f.occl[word], hits);
        }
      }
    }
  }
  __syncthreads();
  st.cyc[3] += (unsigned long long)(clock64() - t0);
}

__global__ void __launch_bounds__(kVisThreads, 3)
vx_visibility_kernel(const VxSlot* __restrict__ slots, const VxVisFrame* __restrict__ frames,
                     const float* __restrict__ endpoints, const VxGridGeom g) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  VisShared& sm = *reinterpret_cast<VisShared*>(smem_raw);
  const VxSlot& slot = slots[blockIdx.x];
  const VxVisFrame vf = frames[blockIdx.x];
  const uint32_t tid = threadIdx.x;
  const uint32_t nvox = (uint32_t)g.nx * (uint32_t)g.ny * (uint32_t)g.nz;
  const uint32_t words = (nvox + 31u) >> 5;

  Frame f;
  f.g = g;
  f.nz_shift = ((g.nz & (g.nz - 1)) == 0) ? (31 - __clz(g.nz)) : -1;
  f.occ = slot.occ;
  f.memo = slot.memo;
  f.rayhit = slot.rayhit;
  f.occl = slot.occl;
  {
    const uint32_t cx = (uint32_t)(g.nx + 3) >> 2, cy = (uint32_t)(g.ny + 3) >> 2, cz = (uint32_t)(g.nz + 1) >> 1;
    if ((cx * cy * cz + 31u) / 32u <= kCoarseWordsMax) {
      f.cy = (int32_t)cy;
      f.cz = (int32_t)cz;
    } else {
      f.cy = f.cz = 0;  // grid too large for the shared-memory index: test the L2-resident bitmap directly
    }
  }
  if (f.cz > 0) build_sparse_occupancy(f, sm);

  Stats st;
  const int32_t shells = vx_num_shells(g.nx, g.ny);
  int32_t cols_per_wave = kWaveCells / g.nz;
  if (cols_per_wave < 1) cols_per_wave = 1;

  for (uint32_t w = tid; w < words; w += kVisThreads) f.occl[w] = 0u;

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
    f.epoch = (pass % kMaxEpoch) + 1u;
    if (pass == 0) {
      for (uint32_t x = tid; x < nvox; x += kVisThreads) f.memo[x] = 0u;
    } else if (pass % kMaxEpoch == 0) {  // epoch counter restarts: forget every memo but keep the dead cells
      for (uint32_t x = tid; x < nvox; x += kVisThreads)
        if (__ldcg(f.memo + x) != kDead) f.memo[x] = 0u;
    }
    __syncthreads();

    uint32_t tau = 0;
    for (int32_t o = 0; o < shells; ++o) {
      for (int32_t face = 0; face < 3; ++face) {
        const int32_t outer = face == 0 ? g.ny : g.nx - o - 1;
        Wave w;
        w.face = face;
        w.fixed = face == 0 ? g.nx - 1 - o : (face == 1 ? o : g.ny - 1 - o);
        for (int32_t a = 0; a < outer; a += cols_per_wave) {
          w.a_begin = a;
          w.ncols = outer - a < cols_per_wave ? outer - a : cols_per_wave;
          w.tau_base = tau;
          run_wave(f, w, sm, st);
          tau += (uint32_t)w.ncols * (uint32_t)g.nz;
        }
      }
    }
    if (pass == 0) {  // invalid_ = occlusions_ (VoxelGrid.cpp:169): what is not occluded is dead from here on
      for (uint32_t x = tid; x < nvox; x += kVisThreads)
        if (!((__ldcg(f.occl + (x >> 5)) >> (x & 31u)) & 1u)) f.memo[x] = kDead;
      __syncthreads();
    }
  }

  // alive bitmap for the packer: invalid_ != -1
  for (uint32_t x0 = 0; x0 < words * 32u; x0 += kVisThreads) {
    const uint32_t x = x0 + tid;
    const bool alive = x < nvox && __ldcg(f.memo + x) != kDead;
    const uint32_t word = __ballot_sync(0xFFFFFFFFu, alive);
    if ((tid & 31u) == 0 && (x >> 5) < words) slot.alive[x >> 5] = word;
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
  if (tid == 0) {
    slot.counters[2] = st.waves;
    slot.counters[3] = st.resolve_waves;
    slot.counters[4] = st.rounds;
    for (int k = 0; k < 4; ++k) slot.counters[5 + k] = st.cyc[k];
  }
}

}  // namespace

cudaError_t vx_visibility_prepare() {
  return cudaFuncSetAttribute(vx_visibility_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(VisShared));
}

cudaError_t vx_launch_visibility(const VxSlot* slots, const VxVisFrame* frames, uint32_t n_slots,
                                 const float* endpoints, VxGridGeom geom, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  vx_visibility_kernel<<<n_slots, kVisThreads, sizeof(VisShared), stream>>>(slots, frames, endpoints, geom);
  return cudaGetLastError();
}

uint64_t vx_visibility_visits(VxGridGeom g) {
  uint64_t v = 0;
  const int32_t shells = vx_num_shells(g.nx, g.ny);
  for (int32_t o = 0; o < shells; ++o) v += (uint64_t)g.nz * ((uint64_t)g.ny + 2ull * (uint64_t)(g.nx - o - 1));
  return v;
}

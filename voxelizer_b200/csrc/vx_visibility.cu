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
//   * memo[L] = (epoch << 26) | (tau_ray << 1) | hit, merged with atomicMax: the value a visit reads
//     is the ray with the largest tau that crossed the cell = the latest writer.  `epoch` = pass
//     number, so the reference's "fill occludedBy_ with -2" costs nothing.
//   * a WAVE = a tau-contiguous run of whole columns of one face (<= kWaveCells cells).  One CTA owns
//     one frame and runs its waves in order; inside a wave
//       A  snapshot the memo of every cell; candidates = visited cells nothing has crossed yet;
//       B  decide which candidates cast: c casts iff no CASTING candidate with smaller tau crosses c
//          with the in-face prefix of its ray (rays are monotone per axis, so they never re-enter a
//          face).  Monotone fix-point over the undecided candidates, all in shared memory;
//       C  casters trace their rays in parallel (atomicMax into memo; cells of the same wave that
//          are visited at or after the caster also merge into the snapshot);
//       D  apply the visit rule from the snapshot: occlusions_[c] = memo (last visit wins) or
//          invalid_[c] = min(invalid_[c], memo).
//   * only booleans are observable (isOccluded / isInvalid, VoxelGrid.cpp:65-73): occluded = hit,
//     alive = hit at every visit of every pass.
// tests/model/wave_model.cpp is the CPU model of exactly this formulation.
//
// Latency-bound, not HBM-bound: dependent reads of one occupancy bit per DDA step (read-only for
// the whole kernel -> L1 resident via ld.global.nc) and scattered 4-byte atomicMax updates in L2.
#include "vx_launch.h"

namespace {

constexpr int kVisThreads = 512;
constexpr int kWaveCells = 4096;
constexpr uint32_t kSkip = 0xFFFFFFFFu;
constexpr uint32_t kInf = 0xFFFFFFFFu;
constexpr uint32_t kEpochShift = 26;
constexpr uint32_t kMaxEpoch = 63;

struct VisShared {
  uint32_t snap[kWaveCells];      // 0 unset | kSkip | (rank << 1 | hit); rank 1 = earlier wave, q + 2 = in-wave caster q
  uint32_t cov_und[kWaveCells];   // min tau (local) of an UNDECIDED candidate crossing the cell
  uint32_t cov_cast[kWaveCells];  // min tau (local) of a DECIDED caster crossing the cell
  uint16_t cand[kWaveCells];      // candidate list (local indices)
  uint8_t status[kWaveCells];     // per candidate slot: 0 undecided, 1 casts, 2 does not
  uint32_t n_cand;
  uint32_t any_forward;
  uint32_t remaining;
};

struct Wave {
  int32_t face;     // 0: i fixed, outer j; 1/2: j fixed, outer i
  int32_t fixed;    // the fixed coordinate
  int32_t a_begin;  // first outer coordinate
  int32_t ncols;
  uint32_t tau_base;
};

struct Frame {
  VxGridGeom g;
  int32_t nz_shift;  // log2(nz) if nz is a power of two, else -1
  const uint32_t* __restrict__ occ;
  uint32_t* memo;
  uint32_t* alive;
  uint32_t* occl;
  float e0, e1, e2;
  uint32_t epoch;
  bool inv;
};

__device__ __forceinline__ uint32_t lin(const VxGridGeom& g, int32_t i, int32_t j, int32_t k) {
  return ((uint32_t)i * (uint32_t)g.ny + (uint32_t)j) * (uint32_t)g.nz + (uint32_t)k;
}

__device__ __forceinline__ bool occupied(const Frame& f, uint32_t L) { return (__ldg(f.occ + (L >> 5)) >> (L & 31u)) & 1u; }

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
// face, hits an occupied cell or terminates: atomicMin(target[cell], qc).  Returns #cells marked.
__device__ uint32_t prefix_mark(const Frame& f, const Wave& w, uint32_t qc, uint32_t* target) {
  int32_t i, j, k;
  cell_of(f, w, qc, i, j, k);
  if (occupied(f, lin(f.g, i, j, k))) return 0;  // an occupied start cell traverses nothing
  VxRay r;
  r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
  const int32_t normal = w.face == 0 ? 0 : 1;
  uint32_t marked = 0;
  for (;;) {
    if (r.next_axis() == normal) break;
    if (!r.advance()) break;
    if (!r.inside(f.g)) break;
    if (occupied(f, lin(f.g, r.i, r.j, r.k))) break;
    const int32_t q = local_of(f, w, r.i, r.j, r.k);
    if (q > (int32_t)qc) {
      atomicMin(&target[q], qc);
      ++marked;
    }
  }
  return marked;
}

// Phase C for one caster.
__device__ void trace(const Frame& f, const Wave& w, VisShared& sm, uint32_t qc, unsigned long long& steps) {
  int32_t i, j, k;
  cell_of(f, w, qc, i, j, k);
  const uint32_t enc0 = (f.epoch << kEpochShift) | ((w.tau_base + qc) << 1);
  const uint32_t senc0 = (qc + 2u) << 1;
  const uint32_t Lc = lin(f.g, i, j, k);
  if (occupied(f, Lc)) {  // returns its own index; the caller's assignment memoises it (VoxelGrid.cpp:112)
    atomicMax(&f.memo[Lc], enc0 | 1u);
    atomicMax(&sm.snap[qc], senc0 | 1u);
    return;
  }
  VxRay r;
  r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
  uint32_t len = 0;
  bool hit = false;
  for (;;) {
    if (!r.inside(f.g)) break;
    const uint32_t L = lin(f.g, r.i, r.j, r.k);
    if (occupied(f, L)) {
      hit = true;
      break;
    }
    atomicMax(&f.memo[L], enc0);
    const int32_t q = local_of(f, w, r.i, r.j, r.k);
    if (q >= (int32_t)qc) atomicMax(&sm.snap[q], senc0);
    ++len;
    if (!r.advance()) break;
  }
  steps += len;
  if (hit) {  // the result is only known at the end: upgrade the crossed cells to "hit"
    r.begin(f.g, i, j, k, f.e0, f.e1, f.e2);
    for (uint32_t s = 0; s < len; ++s) {
      const uint32_t L = lin(f.g, r.i, r.j, r.k);
      atomicMax(&f.memo[L], enc0 | 1u);
      const int32_t q = local_of(f, w, r.i, r.j, r.k);
      if (q >= (int32_t)qc) atomicMax(&sm.snap[q], senc0 | 1u);
      r.advance();
    }
  }
}

__device__ void run_wave(const Frame& f, const Wave& w, VisShared& sm, unsigned long long& rays, unsigned long long& steps) {
  const uint32_t tid = threadIdx.x;
  const uint32_t lane = tid & 31u;
  const uint32_t n = (uint32_t)w.ncols * (uint32_t)f.g.nz;

  if (tid == 0) {
    sm.n_cand = 0;
    sm.any_forward = 0;
    sm.remaining = 0;
  }
  __syncthreads();

  // ---- A: snapshot + candidates
  for (uint32_t q0 = 0; q0 < n; q0 += kVisThreads) {
    const uint32_t q = q0 + tid;
    bool is_cand = false;
    if (q < n) {
      int32_t i, j, k;
      cell_of(f, w, q, i, j, k);
      const uint32_t L = lin(f.g, i, j, k);
      uint32_t s;
      if (f.inv && !((__ldcg(f.alive + (L >> 5)) >> (L & 31u)) & 1u)) {
        s = kSkip;  // invalid_ == -1: "just skip" (VoxelGrid.cpp:187)
      } else {
        const uint32_t m = __ldcg(f.memo + L);
        if ((m >> kEpochShift) == f.epoch) {
          s = 2u | (m & 1u);
        } else {
          s = 0u;
          is_cand = true;
        }
      }
      sm.snap[q] = s;
      sm.cov_und[q] = kInf;
      sm.cov_cast[q] = kInf;
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
  if (n_cand == 0 && !f.inv) {
    // nothing to cast, but the occlusion visit still assigns (falls through to D)
  }

  // ---- B: which candidates cast?
  if (n_cand > 0) {
    for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
      sm.status[p] = 0;
      if (prefix_mark(f, w, sm.cand[p], sm.cov_und)) sm.any_forward = 1;
    }
    __syncthreads();
    if (sm.any_forward) {
      for (;;) {
        // decide what can be decided from cov_cast (decided casters) and cov_und (undecided ones)
        for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
          if (sm.status[p] != 0) continue;
          const uint32_t qc = sm.cand[p];
          if (sm.cov_cast[qc] < qc) {
            sm.status[p] = 2;
          } else if (sm.cov_und[qc] >= qc) {
            sm.status[p] = 1;
            prefix_mark(f, w, qc, sm.cov_cast);
          } else {
            sm.remaining = 1;
          }
        }
        __syncthreads();
        if (!sm.remaining) break;
        __syncthreads();
        if (tid == 0) sm.remaining = 0;
        for (uint32_t q = tid; q < n; q += kVisThreads) sm.cov_und[q] = kInf;
        __syncthreads();
        for (uint32_t p = tid; p < n_cand; p += kVisThreads)
          if (sm.status[p] == 0) prefix_mark(f, w, sm.cand[p], sm.cov_und);
        __syncthreads();
      }
    } else {
      for (uint32_t p = tid; p < n_cand; p += kVisThreads) sm.status[p] = 1;
      __syncthreads();
    }

    // ---- C: trace
    for (uint32_t p = tid; p < n_cand; p += kVisThreads) {
      if (sm.status[p] != 1) continue;
      ++rays;
      trace(f, w, sm, sm.cand[p], steps);
    }
    __syncthreads();
  }

  // ---- D: apply the visit
  for (uint32_t q0 = 0; q0 < n; q0 += kVisThreads) {
    const uint32_t q = q0 + tid;
    bool act = false, hit = false;
    uint32_t L = 0;
    if (q < n) {
      const uint32_t s = sm.snap[q];
      if (s != kSkip) {
        int32_t i, j, k;
        cell_of(f, w, q, i, j, k);
        L = lin(f.g, i, j, k);
        hit = (s & 1u) != 0;
        act = f.inv ? !hit : true;  // invalid pass: only a miss changes state (alive -> dead)
      }
    }
    const uint32_t active = __ballot_sync(0xFFFFFFFFu, act);
    if (act) {
      const uint32_t word = L >> 5;
      const uint32_t bit = 1u << (L & 31u);
      const uint32_t peers = __match_any_sync(active, word);
      const uint32_t visited = __reduce_or_sync(peers, bit);
      const uint32_t hits = __reduce_or_sync(peers, hit ? bit : 0u);
      if ((uint32_t)(__ffs(peers) - 1) == lane) {
        if (f.inv) {
          atomicAnd(&f.alive[word], ~visited);
        } else {
          if (visited != hits) atomicAnd(&f.occl[word], ~(visited & ~hits));
          if (hits) atomicOr(&f.occl[word], hits);
        }
      }
    }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(kVisThreads, 2)
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
  f.alive = slot.alive;
  f.occl = slot.occl;

  unsigned long long rays = 0, steps = 0;
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
    if (pass % kMaxEpoch == 0) {  // epoch counter (re)starts: forget every memo
      for (uint32_t x = tid; x < nvox; x += kVisThreads) f.memo[x] = 0u;
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
          run_wave(f, w, sm, rays, steps);
          tau += (uint32_t)w.ncols * (uint32_t)g.nz;
        }
      }
    }
    if (pass == 0) {  // invalid_ = occlusions_ (VoxelGrid.cpp:169)
      for (uint32_t x = tid; x < words; x += kVisThreads) f.alive[x] = __ldcg(f.occl + x);
      __syncthreads();
    }
  }

  // statistics
  for (int off = 16; off > 0; off >>= 1) {
    rays += __shfl_down_sync(0xFFFFFFFFu, rays, off);
    steps += __shfl_down_sync(0xFFFFFFFFu, steps, off);
  }
  if ((tid & 31u) == 0) {
    atomicAdd(&slot.counters[0], rays);
    atomicAdd(&slot.counters[1], steps);
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

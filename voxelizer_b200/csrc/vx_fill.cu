// K1 + K2 (fill), K3 (vote), K5 (emit / pack): the HBM-side passes of the gen_data hot path.
//
//   fill  = fillVoxelGrid + VoxelGrid::insert     voxelize_utils.cpp:173-203, VoxelGrid.cpp:45-63
//   vote  = majority label of saveVoxelGrid        voxelize_utils.cpp:237-245
//   emit  = .label / .bin output + pack()          voxelize_utils.cpp:205-216, 257-294
//   pack  = .occluded / .invalid output            voxelize_utils.cpp:267-283, VoxelGrid.cpp:65-73
//
// All fp32 arithmetic that decides which voxel a point falls into is written with the
// round-to-nearest intrinsics (__fmul_rn, __fadd_rn, __fdiv_rn, __fsqrt_rn): the compiler never
// contracts those into FMAs, which is what reproduces the reference's SSE2 (FMA-free) Eigen code.
//
// The reference re-reads a frame's whole window for every frame (180 MB of points per frame).  Here the
// fill is organised by SCAN: a CTA loads a tile of a scan ONCE into registers, applies the
// frame-independent filters once, and then walks the frames of the launch whose window holds that scan
// (~14 at stride 5, 71 in the dense configuration), transforming with that frame's matrix and inserting
// into that frame's histogram.  The histograms of the frames a scan touches (~2 MB each) stay in L2, so
// the DRAM traffic of the pass is a fraction of its ALGORITHMIC bytes (20 B per window point per frame).
#include "vx_launch.h"

namespace {

constexpr int kFillThreads = 256;
constexpr int kFillPointsPerThread = 2;
constexpr int kFillTile = kFillThreads * kFillPointsPerThread;
constexpr int kFillUsesPerChunk = 16;  // frames staged in shared memory at a time

// slot of key (L << 16 | label) + 1: two 32-bit multiplicative mixes (L and label are < 2^31 and < 2^16)
__device__ __forceinline__ uint32_t hash_slot(unsigned long long key1, uint32_t log2_cap) {
  const uint32_t lo = (uint32_t)key1, hi = (uint32_t)(key1 >> 16);
  return ((hi * 0x9E3779B1u) ^ (lo * 0x85EBCA77u)) >> (32u - log2_cap);
}

// floorf(__fdiv_rn(v, res)) for the three coordinates without the divisions in the common case.  q = v * rcp
// (rcp = RN(1/res)) is within 1.5 * 2^-23 relative of the correctly rounded quotient, so when q is further than
// 2^-21 |q| from both neighbouring integers the two have the same floor; otherwise (about 1 point in 10^4), and
// for NaN / infinity (every comparison false), the IEEE divisions decide.
__device__ __forceinline__ void floor_div3(float vx, float vy, float vz, float res, float rcp, float& fx, float& fy, float& fz) {
  const float qx = __fmul_rn(vx, rcp), qy = __fmul_rn(vy, rcp), qz = __fmul_rn(vz, rcp);
  fx = floorf(qx);
  fy = floorf(qy);
  fz = floorf(qz);
  // |d - 0.5| < 0.5 - margin  <=>  d > margin and 1 - d > margin, with d = q - floor(q) in [0, 1) (exact)
  const float kMargin = 4.76837158e-7f;  // 2^-21
  const bool safe = fabsf((qx - fx) - 0.5f) < 0.5f - fabsf(qx) * kMargin && fabsf((qy - fy) - 0.5f) < 0.5f - fabsf(qy) * kMargin &&
                    fabsf((qz - fz) - 0.5f) < 0.5f - fabsf(qz) * kMargin;
  if (!safe) {  // about 1 point in 10^4, and every NaN / infinity: the IEEE divisions decide
    fx = floorf(__fdiv_rn(vx, res));
    fy = floorf(__fdiv_rn(vy, res));
    fz = floorf(__fdiv_rn(vz, res));
  }
}

// (voxel,label) -> count.  `seen` = the slot's key as loaded earlier by the caller (possibly stale: a
// stale 0 is corrected by the CAS).  Keys go 0 -> key exactly once per frame.
__device__ __forceinline__ void table_add(unsigned long long* keys, uint32_t* counts, uint32_t log2_cap, uint32_t* overflow,
                                          unsigned long long key1, uint32_t h, unsigned long long seen, uint32_t n) {
  const uint32_t mask = (1u << log2_cap) - 1u;
  for (uint32_t probe = 0; probe < VX_MAX_PROBES; ++probe) {
    if (seen == 0ull) {
      seen = atomicCAS(&keys[h], 0ull, key1);
      if (seen == 0ull) seen = key1;
    }
    if (seen == key1) {
      vx_red_add(&counts[h], n);
      return;
    }
    h = (h + 1u) & mask;
    seen = __ldcg(&keys[h]);
  }
  atomicExch(overflow, 1u);
}

struct FillUse {  // one frame a scan goes into, staged in shared memory
  float m[12];    // rows 0..2 of ap, column-major: m[c*3 + r]
  unsigned long long* keys_t;
  uint32_t* counts_t;
  uint32_t* ovf_t;
  unsigned long long* keys_i;
  uint32_t* counts_i;
  uint32_t* ovf_i;
  uint32_t log2_t, log2_i;
  uint32_t both;
  uint32_t skip;  // the tile's bounding box, moved into this frame, lies clearly outside the grid
};

__global__ void __launch_bounds__(kFillThreads, 8)
vx_fill_kernel(const VxScanWork* __restrict__ works, const VxScanUse* __restrict__ uses, const float* __restrict__ ap_all,
               const VxSlot* __restrict__ slots, const VxGridGeom g, const VxFilterDev f) {
  __shared__ FillUse su[kFillUsesPerChunk];
  __shared__ float bbox[kFillThreads / 32][6];  // per warp: min x,y,z, max x,y,z of the points that passed the filters
  const VxScanWork w = works[blockIdx.y];
  const uint32_t base = blockIdx.x * kFillTile;
  if (base >= w.n) return;
  const uint32_t lane = threadIdx.x & 31u;

  // ---- the tile, once: points, labels, frame-independent filters (voxelize_utils.cpp:186-198)
  float px[kFillPointsPerThread], py[kFillPointsPerThread], pz[kFillPointsPerThread];
  uint32_t lab[kFillPointsPerThread];
  uint32_t keep0 = 0;
  {
    float4 p[kFillPointsPerThread];
#ifdef VX_FILL_TMA
    // A/B (tools/build_variants.py ...,VX_FILL_TMA=1): the tile staged in shared memory by ONE bulk copy per array (cp.async.bulk +
    // mbarrier, the 1-D TMA path; SASS UBLKCP) instead of one ld.global.cs per point and thread.  Full tiles only (16-byte sizes).
    __shared__ __align__(128) float4 tile_points[kFillTile];
    __shared__ __align__(16) uint32_t tile_labels[kFillTile];
    __shared__ __align__(8) unsigned long long tile_bar;
    const bool bulk = base + kFillTile <= w.n;
    if (bulk) {
      const uint32_t bar = (uint32_t)__cvta_generic_to_shared(&tile_bar);
      if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
      }
      __syncthreads();
      if (threadIdx.x == 0) {
        const uint32_t bytes_p = kFillTile * 16u, bytes_l = kFillTile * 4u;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes_p + bytes_l) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(tile_points)), "l"(w.points + base), "r"(bytes_p), "r"(bar) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                         (uint32_t)__cvta_generic_to_shared(tile_labels)), "l"(w.labels + base), "r"(bytes_l), "r"(bar) : "memory");
      }
      uint32_t done = 0;
      while (!done) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(done) : "r"(bar) : "memory");
      }
    }
#else
    constexpr bool bulk = false;
#endif
#pragma unroll
    for (int u = 0; u < kFillPointsPerThread; ++u) {
      const uint32_t idx = base + u * kFillThreads + threadIdx.x;
      p[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      lab[u] = 0;
      if (idx < w.n) {
#ifdef VX_FILL_TMA
        if (bulk) {
          p[u] = tile_points[u * kFillThreads + threadIdx.x];
          lab[u] = tile_labels[u * kFillThreads + threadIdx.x] & 0xFFFFu;
        } else
#endif
        {
          p[u] = __ldcs(w.points + idx);  // a scan is streamed once per launch: evict-first
          lab[u] = __ldcs(w.labels + idx) & 0xFFFFu;
        }
        keep0 |= 1u << u;
      }
    }
#pragma unroll
    for (int u = 0; u < kFillPointsPerThread; ++u) {
      const float x = p[u].x, y = p[u].y, z = p[u].z;
      px[u] = x;
      py[u] = y;
      pz[u] = z;
      bool keep = (keep0 >> u) & 1u;
      // voxelize_utils.cpp:188  Eigen norm: sqrtf(x*x + (y*y + z*z)); a NaN range passes, as in the reference
      const float range = __fsqrt_rn(__fadd_rn(__fmul_rn(x, x), __fadd_rn(__fmul_rn(y, y), __fmul_rn(z, z))));
      if (range < f.min_range || range > f.max_range) keep = false;
      // voxelize_utils.cpp:190  ego-vehicle box
      if (f.hidecar && x < 3.0f && x > -2.0f && fabsf(y) < 2.0f) keep = false;
      // voxelize_utils.cpp:195-198 (join, then ignore) as a 64 Ki-bit LUT over the masked label
      if (keep && f.has_label_filter && ((__ldg(f.reject_bits + (lab[u] >> 5)) >> (lab[u] & 31u)) & 1u)) keep = false;
      if (!keep) keep0 &= ~(1u << u);
    }
  }
  // bounding box of the surviving points (a frame whose grid the box cannot touch is skipped as a whole)
  {
    float lo[3] = {INFINITY, INFINITY, INFINITY}, hi[3] = {-INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (int u = 0; u < kFillPointsPerThread; ++u)
      if ((keep0 >> u) & 1u) {
        lo[0] = fminf(lo[0], px[u]); hi[0] = fmaxf(hi[0], px[u]);
        lo[1] = fminf(lo[1], py[u]); hi[1] = fmaxf(hi[1], py[u]);
        lo[2] = fminf(lo[2], pz[u]); hi[2] = fmaxf(hi[2], pz[u]);
      }
#pragma unroll
    for (int a = 0; a < 3; ++a)
#pragma unroll
      for (int off = 16; off > 0; off >>= 1) {
        lo[a] = fminf(lo[a], __shfl_xor_sync(0xFFFFFFFFu, lo[a], off));
        hi[a] = fmaxf(hi[a], __shfl_xor_sync(0xFFFFFFFFu, hi[a], off));
      }
    if (lane == 0) {
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        bbox[threadIdx.x >> 5][a] = lo[a];
        bbox[threadIdx.x >> 5][3 + a] = hi[a];
      }
    }
  }
  if (__syncthreads_or(keep0 != 0u) == 0) return;  // nothing of this tile survives the filters

  const float fnx = (float)g.nx, fny = (float)g.ny, fnz = (float)g.nz;
  const float rcp = __frcp_rn(g.res);
  for (uint32_t q0 = 0; q0 < w.use_count; q0 += kFillUsesPerChunk) {
    const uint32_t nq = min((uint32_t)kFillUsesPerChunk, w.use_count - q0);
    __syncthreads();  // the previous chunk has been consumed
    if (threadIdx.x < nq) {
      const VxScanUse use = uses[w.use_begin + q0 + threadIdx.x];
      const VxSlot& s = slots[use.slot];
      const float* m = ap_all + 16u * use.ap_index;  // column-major 4x4: m[c*4 + r]
      FillUse& d = su[threadIdx.x];
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int r = 0; r < 3; ++r) d.m[c * 3 + r] = m[c * 4 + r];
      d.keys_t = s.target.keys;
      d.counts_t = s.target.counts;
      d.ovf_t = s.target.overflow;
      d.log2_t = s.target.log2_cap;
      d.keys_i = s.input.keys;
      d.counts_i = s.input.counts;
      d.ovf_i = s.input.overflow;
      d.log2_i = s.input.log2_cap;
      d.both = use.both;
      // Where can the tile land in this frame?  centre' +- |R| * half-extent bounds every transformed point (NaN
      // coordinates never enter fminf / fmaxf and are rejected per point anyway).  One metre of slack plus 1e-5 of the
      // operands' magnitude dwarfs any rounding difference between this bound and the per-point arithmetic.
      float lo[3], hi[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        lo[a] = bbox[0][a];
        hi[a] = bbox[0][3 + a];
#pragma unroll
        for (int wv = 1; wv < kFillThreads / 32; ++wv) {
          lo[a] = fminf(lo[a], bbox[wv][a]);
          hi[a] = fmaxf(hi[a], bbox[wv][3 + a]);
        }
      }
      const float cx = 0.5f * (lo[0] + hi[0]), cy = 0.5f * (lo[1] + hi[1]), cz = 0.5f * (lo[2] + hi[2]);
      const float hx = 0.5f * (hi[0] - lo[0]), hy = 0.5f * (hi[1] - lo[1]), hz = 0.5f * (hi[2] - lo[2]);
      const float gsize[3] = {(float)g.nx * g.res, (float)g.ny * g.res, (float)g.nz * g.res};
      bool outside = false;
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        const float centre = m[0 * 4 + r] * cx + m[1 * 4 + r] * cy + m[2 * 4 + r] * cz + m[3 * 4 + r];
        // the slack grows with the magnitude of the operands: beyond ~1.6e7 one ulp exceeds a fixed metre (a config may
        // allow such points: max range 1e9), and a tile of near points that holds one far point must not be dropped
        const float mag = fabsf(m[0 * 4 + r]) * (fabsf(cx) + hx) + fabsf(m[1 * 4 + r]) * (fabsf(cy) + hy) +
                          fabsf(m[2 * 4 + r]) * (fabsf(cz) + hz) + fabsf(m[3 * 4 + r]);
        const float reach = fabsf(m[0 * 4 + r]) * hx + fabsf(m[1 * 4 + r]) * hy + fabsf(m[2 * 4 + r]) * hz + 1.0f + 1.0e-5f * mag;
        outside |= centre + reach < g.off[r] || centre - reach > g.off[r] + gsize[r];
      }
      d.skip = outside ? 1u : 0u;
    }
    __syncthreads();

    for (uint32_t q = 0; q < nq; ++q) {
      const FillUse& d = su[q];
      if (d.skip) continue;
      const float m00 = d.m[0], m10 = d.m[1], m20 = d.m[2];
      const float m01 = d.m[3], m11 = d.m[4], m21 = d.m[5];
      const float m02 = d.m[6], m12 = d.m[7], m22 = d.m[8];
      const float m03 = d.m[9], m13 = d.m[10], m23 = d.m[11];
      const uint32_t log2_t = d.log2_t;

      // ---- K1: key of every point of the tile in this frame
      unsigned long long key1[kFillPointsPerThread];  // (L << 16 | label) + 1, 0 = not inserted
#pragma unroll
      for (int u = 0; u < kFillPointsPerThread; ++u) {
        const float x = px[u], y = py[u], z = pz[u];
        // voxelize_utils.cpp:193  ((m_r0*x + m_r1*y) + m_r2*z) + m_r3
        const float tx = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m00, x), __fmul_rn(m01, y)), __fmul_rn(m02, z)), m03);
        const float ty = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m10, x), __fmul_rn(m11, y)), __fmul_rn(m12, z)), m13);
        const float tz = __fadd_rn(__fadd_rn(__fadd_rn(__fmul_rn(m20, x), __fmul_rn(m21, y)), __fmul_rn(m22, z)), m23);
        // VoxelGrid.cpp:46-53  floor((p - offset) / resolution), IEEE division; NaN / out of range rejected
        float fi, fj, fk;
        floor_div3(__fsub_rn(tx, g.off[0]), __fsub_rn(ty, g.off[1]), __fsub_rn(tz, g.off[2]), g.res, rcp, fi, fj, fk);
        const bool in = ((keep0 >> u) & 1u) && (fi >= 0.0f) && (fi < fnx) && (fj >= 0.0f) && (fj < fny) && (fk >= 0.0f) && (fk < fnz);
        const uint32_t L = ((uint32_t)fi * (uint32_t)g.ny + (uint32_t)fj) * (uint32_t)g.nz + (uint32_t)fk;
        key1[u] = in ? ((((unsigned long long)L << 16) | (unsigned long long)lab[u]) + 1ull) : 0ull;
      }

      // ---- K2: one table update per RUN of equal keys in a warp row (scan order is beam-major /
      // azimuth-minor: neighbouring lanes mostly share a voxel).  All probes of a thread are in flight together.
      uint32_t runlen[kFillPointsPerThread], slot_t[kFillPointsPerThread];
      unsigned long long seen_t[kFillPointsPerThread];
#pragma unroll
      for (int u = 0; u < kFillPointsPerThread; ++u) {
        const unsigned long long prev = __shfl_up_sync(0xFFFFFFFFu, key1[u], 1);
        const bool lead = key1[u] != 0ull && (lane == 0u || prev != key1[u]);
        const uint32_t starts = __ballot_sync(0xFFFFFFFFu, lead || key1[u] == 0ull);  // a new run, or a gap, begins here
        const uint32_t after = lane == 31u ? 0u : (starts >> (lane + 1u));
        runlen[u] = lead ? (after ? (uint32_t)__ffs(after) : 32u - lane) : 0u;
        slot_t[u] = hash_slot(key1[u], log2_t);
        seen_t[u] = 0ull;
        if (runlen[u]) seen_t[u] = __ldcg(d.keys_t + slot_t[u]);
      }
#pragma unroll
      for (int u = 0; u < kFillPointsPerThread; ++u)
        if (runlen[u]) table_add(d.keys_t, d.counts_t, log2_t, d.ovf_t, key1[u], slot_t[u], seen_t[u], runlen[u]);
      if (d.both) {  // prior scans also fill the input grid (gen_data.cpp:96-97)
#pragma unroll
        for (int u = 0; u < kFillPointsPerThread; ++u)
          if (runlen[u]) {
            const uint32_t h = hash_slot(key1[u], d.log2_i);
            table_add(d.keys_i, d.counts_i, d.log2_i, d.ovf_i, key1[u], h, __ldcg(d.keys_i + h), runlen[u]);
          }
      }
    }
  }
}

// bit b of a 32-voxel word = voxel L0 + b.  pack() puts element 8n in bit 7 of byte n: reverse the
// bits of every byte = reverse the word, then swap its bytes back.
__device__ __forceinline__ uint32_t msb_first(uint32_t word) { return __byte_perm(__brev(word), 0u, 0x0123u); }

// K3a.  Zeroes the dense per-frame outputs the vote scatters into (.label, occupancy, .bin).
__global__ void __launch_bounds__(256) vx_clear_kernel(const VxSlot* __restrict__ slots, uint64_t nvox, uint32_t words) {
  const VxSlot& s = slots[blockIdx.y];
  const uint32_t stride = gridDim.x * blockDim.x, t0 = blockIdx.x * blockDim.x + threadIdx.x;
  uint4* lab4 = reinterpret_cast<uint4*>(s.out_label);  // arenas are 256-byte aligned
  const uint64_t n4 = nvox / 8;
  for (uint64_t x = t0; x < n4; x += stride) lab4[x] = make_uint4(0u, 0u, 0u, 0u);
  for (uint64_t x = n4 * 8 + t0; x < nvox; x += stride) s.out_label[x] = 0;
  for (uint32_t x = t0; x < words; x += stride) {
    s.occ[x] = 0u;
    s.out_bin[x] = 0u;
  }
}

constexpr int kScanBatch = 16; // histogram slots a thread has in flight (the scans are latency-bound otherwise)

// K3b.  best[L] = max over the labels of voxel L of (count << 16 | 0xFFFF - label): largest count, ties ->
// smallest label -- the result of scanning the reference's std::map in ascending label order with a
// strict '>' (voxelize_utils.cpp:237-245).  One fire-and-forget RED per histogram slot.
__global__ void __launch_bounds__(256) vx_vote_kernel(const VxSlot* __restrict__ slots) {
  const VxSlot& s = slots[blockIdx.y >> 1];
  const VxTable& t = (blockIdx.y & 1u) ? s.target : s.input;
  const uint32_t cap = 1u << t.log2_cap;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t h0 = blockIdx.x * blockDim.x + threadIdx.x; h0 < cap; h0 += stride * kScanBatch) {
    unsigned long long k1[kScanBatch];
    uint32_t c[kScanBatch];
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      const uint32_t h = h0 + (uint32_t)u * stride;
      k1[u] = h < cap ? __ldcg(&t.keys[h]) : 0ull;
      c[u] = h < cap ? __ldcg(&t.counts[h]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      if (k1[u] == 0ull) continue;
      const unsigned long long key = k1[u] - 1ull;
      vx_red_max64(&t.best[key >> 16], ((unsigned long long)c[u] << 16) | (0xFFFFull - (key & 0xFFFFull)));
    }
  }
}

// K5a.  The slot that holds its voxel's winner writes .label (uint16), the occupancy bit (target grid) or the
// .bin bit (input grid, majority label > 0, MSB-first) and resets best[L]; every slot is freed (unless kept).
__global__ void __launch_bounds__(256) vx_emit_kernel(const VxSlot* __restrict__ slots, int keep_tables) {
  const VxSlot& s = slots[blockIdx.y >> 1];
  const bool is_target = (blockIdx.y & 1u) != 0;
  const VxTable& t = is_target ? s.target : s.input;
  const uint32_t cap = 1u << t.log2_cap;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t h0 = blockIdx.x * blockDim.x + threadIdx.x; h0 < cap; h0 += stride * kScanBatch) {
    unsigned long long k1[kScanBatch], b[kScanBatch];
    uint32_t c[kScanBatch];
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      const uint32_t h = h0 + (uint32_t)u * stride;
      k1[u] = h < cap ? __ldcs(&t.keys[h]) : 0ull;
      c[u] = h < cap ? __ldcs(&t.counts[h]) : 0u;
    }
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      b[u] = 0ull;
      if (k1[u] != 0ull) {
        b[u] = __ldcg(&t.best[(k1[u] - 1ull) >> 16]);
        if (!keep_tables) {
          const uint32_t h = h0 + (uint32_t)u * stride;
          t.keys[h] = 0ull;
          t.counts[h] = 0u;
        }
      }
    }
#pragma unroll
    for (int u = 0; u < kScanBatch; ++u) {
      if (k1[u] == 0ull) continue;
      const unsigned long long key = k1[u] - 1ull;
      const uint32_t L = (uint32_t)(key >> 16);
      const uint32_t label = (uint32_t)(key & 0xFFFFull);
      if (b[u] != (((unsigned long long)c[u] << 16) | (0xFFFFull - label))) continue;  // (count, label) pairs of a voxel are distinct
      t.best[L] = 0ull;
      if (is_target) {
        s.out_label[L] = (uint16_t)label;
        vx_red_or(&s.occ[L >> 5], 1u << (L & 31u));  // occupied = count > 0, whatever the label (VoxelGrid.cpp:45-63)
      } else if (label > 0u) {
        const uint32_t bb = L & 31u;
        vx_red_or(&s.out_bin[L >> 5], 1u << ((bb & 24u) | (7u - (bb & 7u))));  // element 8n + e -> byte n, bit 7 - e
      }
    }
  }
}

// Frees tables that were kept for debugging.
__global__ void __launch_bounds__(256) vx_table_reset_kernel(const VxSlot* __restrict__ slots) {
  const VxSlot& s = slots[blockIdx.y >> 1];
  const VxTable& t = (blockIdx.y & 1u) ? s.target : s.input;
  const uint32_t cap = 1u << t.log2_cap;
  for (uint32_t h = blockIdx.x * blockDim.x + threadIdx.x; h < cap; h += gridDim.x * blockDim.x) {
    t.keys[h] = 0ull;
    t.counts[h] = 0u;
  }
}

// K5b.  isOccluded = occlusions_ > -1; isInvalid = invalid_ > -1 && invalid_ != own index, i.e. still
// alive after every pass and not occupied (VoxelGrid.cpp:65-73).
__global__ void __launch_bounds__(256) vx_pack_kernel(const VxSlot* __restrict__ slots, uint32_t words) {
  const VxSlot& s = slots[blockIdx.y];
  for (uint32_t w = blockIdx.x * blockDim.x + threadIdx.x; w < words; w += gridDim.x * blockDim.x) {
    const uint32_t occ = s.occ[w], occl = s.occl[w], alive = s.alive[w];
    s.out_occluded[w] = msb_first(occl);
    s.out_invalid[w] = msb_first(alive & ~occ);
  }
}

}  // namespace

cudaError_t vx_launch_fill(const VxScanWork* works, uint32_t n_works, uint32_t max_points, const VxScanUse* uses, const float* ap,
                           const VxSlot* slots, VxGridGeom geom, VxFilterDev filter, cudaStream_t stream) {
  if (n_works == 0 || max_points == 0) return cudaSuccess;
  const uint32_t tiles = (max_points + kFillTile - 1) / kFillTile;
  for (uint32_t first = 0; first < n_works; first += 65535u) {
    const uint32_t n = n_works - first < 65535u ? n_works - first : 65535u;
    vx_fill_kernel<<<dim3(tiles, n), kFillThreads, 0, stream>>>(works + first, uses, ap, slots, geom, filter);
  }
  return cudaGetLastError();
}

cudaError_t vx_launch_vote(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  const uint32_t words = (uint32_t)((nvox + 31) / 32);
  vx_clear_kernel<<<dim3(64, n_slots), 256, 0, stream>>>(slots, nvox, words);
  vx_vote_kernel<<<dim3(32, 2 * n_slots), 256, 0, stream>>>(slots);
  return cudaGetLastError();
}

cudaError_t vx_launch_table_reset(const VxSlot* slots, uint32_t n_slots, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  vx_table_reset_kernel<<<dim3(64, 2 * n_slots), 256, 0, stream>>>(slots);
  return cudaGetLastError();
}

cudaError_t vx_launch_emit(const VxSlot* slots, uint32_t n_slots, int keep_tables, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  vx_emit_kernel<<<dim3(32, 2 * n_slots), 256, 0, stream>>>(slots, keep_tables);
  return cudaGetLastError();
}

cudaError_t vx_launch_pack(const VxSlot* slots, uint32_t n_slots, uint64_t nvox, cudaStream_t stream) {
  if (n_slots == 0) return cudaSuccess;
  const uint32_t words = (uint32_t)((nvox + 31) / 32);
  uint32_t blocks = (words + 255) / 256;
  if (blocks > 256u) blocks = 256u;
  vx_pack_kernel<<<dim3(blocks, n_slots), 256, 0, stream>>>(slots, words);
  return cudaGetLastError();
}

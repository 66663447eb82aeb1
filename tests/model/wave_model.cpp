// TEST INFRASTRUCTURE: sequential CPU model of the wave-parallel visibility algorithm that
// voxelizer_b200/csrc/vx_visibility.cu runs on the GPU ("K4").  It shares the product's
// ray arithmetic header (vx_dda.h) and executes the kernel's phases in an ORDER-SCRAMBLED way (casters
// are traced in a pseudo-random order, merges are max-reductions) so that a test passing here shows
// the formulation is independent of thread scheduling.  tests/test_wave_model.py checks it bit for
// bit against the oracle (reference VoxelGrid::updateOcclusions / updateInvalid, VoxelGrid.cpp:98-316).
//
// The algorithm, per pass (one endpoint) -- identical to the kernel:
//   cells are visited in the reference order; visit number tau.  A WAVE is a tau-contiguous run of
//   whole columns of one face of one shell (at most W cells).  Waves are processed in order:
//   A. snapshot: for each cell of the wave read the memo (latest ray that crossed it so far in this
//      pass).  skip = invalid pass and cell no longer alive.  candidate = not skip and memo unset.
//   B. resolve which candidates really cast: c casts iff no CASTING candidate with smaller tau covers
//      c with the part of its ray that lies inside this wave (rays are monotone per axis, so a ray
//      never re-enters the face).  Solved as a monotone fix-point over undecided candidates.
//   C. every caster traces its ray; each traversed cell t gets memo[t] = max(memo[t], (tau,hit));
//      cells of the same wave with tau_t >= tau_caster also merge into the snapshot.
//   D. apply the visit rule from the snapshot: occlusion pass  occluded[c] = hit (last visit wins);
//      invalid pass  alive[c] &= hit.
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../voxelizer_b200/csrc/vx_dda.h"

namespace {

struct Model {
  VxGridGeom g;
  size_t nvox;
  std::vector<uint32_t> occ, alive, occl;  // bitmaps over L = (i*ny + j)*nz + k
  std::vector<uint32_t> memo;              // (epoch << 26) | (tau << 1) | hit
  uint32_t epoch = 0;
  uint32_t wave_cells = 8192;
  uint64_t rng = 1;
  // stats
  uint64_t n_rays = 0, n_steps = 0, n_waves = 0, n_resolve_waves = 0, n_rounds = 0, max_rounds = 0, n_prefix_steps = 0, n_unset = 0;

  static bool bit(const std::vector<uint32_t>& b, size_t L) { return (b[L >> 5] >> (L & 31)) & 1u; }
  static void setbit(std::vector<uint32_t>& b, size_t L, bool v) {
    if (v) b[L >> 5] |= (1u << (L & 31));
    else b[L >> 5] &= ~(1u << (L & 31));
  }
  size_t lin(int32_t i, int32_t j, int32_t k) const { return (size_t(i) * size_t(g.ny) + size_t(j)) * size_t(g.nz) + size_t(k); }

  uint32_t next_rand() {
    rng = rng * 6364136223846793005ull + 1442695040888963407ull;
    return uint32_t(rng >> 33);
  }

  struct Wave {
    int face, o;
    int32_t a_begin, a_end;  // outer coordinate range
    uint32_t tau_base;
    int32_t fixed;  // i for face 0, j for faces 1/2
  };

  void cell_of(const Wave& w, uint32_t q, int32_t& i, int32_t& j, int32_t& k) const {
    const int32_t col = int32_t(q / uint32_t(g.nz));
    k = int32_t(q % uint32_t(g.nz));
    if (w.face == 0) {
      i = w.fixed;
      j = w.a_begin + col;
    } else {
      i = w.a_begin + col;
      j = w.fixed;
    }
  }
  // local index of (i,j,k) inside wave w, or -1
  int32_t local_of(const Wave& w, int32_t i, int32_t j, int32_t k) const {
    if (w.face == 0) {
      if (i != w.fixed || j < w.a_begin || j >= w.a_end) return -1;
      return (j - w.a_begin) * g.nz + k;
    }
    if (j != w.fixed || i < w.a_begin || i >= w.a_end) return -1;
    return (i - w.a_begin) * g.nz + k;
  }

  // in-wave cells (tau larger than the caster's) crossed by candidate c's ray before it leaves the
  // face, hits, or terminates
  template <class F>
  void prefix_walk(const Wave& w, uint32_t qc, const float* e, F f) {
    int32_t i, j, k;
    cell_of(w, qc, i, j, k);
    if (bit(occ, lin(i, j, k))) return;  // occupied start: nothing traversed
    VxRay r;
    r.begin(g, i, j, k, e[0], e[1], e[2]);
    for (;;) {
      const int32_t a = r.next_axis();
      const bool leaves_face = (w.face == 0) ? (a == 0) : (a == 1);
      if (leaves_face) return;
      if (!r.advance()) return;
      if (!r.inside(g)) return;
      ++n_prefix_steps;
      if (bit(occ, lin(r.i, r.j, r.k))) return;
      const int32_t q = local_of(w, r.i, r.j, r.k);
      if (q > int32_t(qc)) f(uint32_t(q));
    }
  }

  void run_wave(const Wave& w, const float* e, bool inv) {
    const uint32_t n = uint32_t(w.a_end - w.a_begin) * uint32_t(g.nz);
    std::vector<uint32_t> snap(n, 0);  // 0 unset; (rank << 1) | hit, rank 1 = earlier waves, q+2 = in-wave caster q
    std::vector<uint8_t> skip(n, 0);
    std::vector<uint32_t> cand;
    // ---- A
    for (uint32_t q = 0; q < n; ++q) {
      int32_t i, j, k;
      cell_of(w, q, i, j, k);
      const size_t L = lin(i, j, k);
      if (inv && !bit(alive, L)) {
        skip[q] = 1;
        continue;
      }
      const uint32_t m = memo[L];
      if ((m >> 26) == epoch) snap[q] = 2u | (m & 1u);
      else cand.push_back(q);
    }
    ++n_waves;
    // ---- B (monotone fix-point; status 0 undecided, 1 cast, 2 not)
    std::vector<uint8_t> status(n, 0);
    std::vector<uint32_t> casters;
    {
      bool any_forward = false;
      for (uint32_t qc : cand) prefix_walk(w, qc, e, [&](uint32_t) { any_forward = true; });
      if (!any_forward) {
        casters = cand;
      } else {
        ++n_resolve_waves;
        const uint32_t INF = 0xFFFFFFFFu;
        std::vector<uint32_t> cov_cast(n, INF);
        std::vector<uint32_t> undecided = cand;
        uint64_t rounds = 0;
        while (!undecided.empty()) {
          ++rounds;
          std::vector<uint32_t> cov_und(n, INF);
          for (uint32_t qc : undecided) prefix_walk(w, qc, e, [&](uint32_t q) { cov_und[q] = std::min(cov_und[q], qc); });
          std::vector<uint32_t> next;
          std::vector<uint32_t> newly;
          for (uint32_t qc : undecided) {
            if (cov_cast[qc] < qc) status[qc] = 2;
            else if (cov_und[qc] >= qc) {
              status[qc] = 1;
              newly.push_back(qc);
            } else next.push_back(qc);
          }
          for (uint32_t qc : newly) prefix_walk(w, qc, e, [&](uint32_t q) { cov_cast[q] = std::min(cov_cast[q], qc); });
          undecided.swap(next);
        }
        n_rounds += rounds;
        max_rounds = std::max(max_rounds, rounds);
        for (uint32_t qc : cand)
          if (status[qc] == 1) casters.push_back(qc);
      }
    }
    // ---- C (scrambled order)
    for (size_t a = casters.size(); a > 1; --a) std::swap(casters[a - 1], casters[next_rand() % a]);
    for (uint32_t qc : casters) {
      int32_t i, j, k;
      cell_of(w, qc, i, j, k);
      const uint32_t tau = w.tau_base + qc;
      ++n_rays;
      bool hit = false;
      uint32_t len = 0;
      if (bit(occ, lin(i, j, k))) {
        hit = true;
      } else {
        VxRay r;
        r.begin(g, i, j, k, e[0], e[1], e[2]);
        for (;;) {
          if (!r.inside(g)) break;
          if (bit(occ, lin(r.i, r.j, r.k))) {
            hit = true;
            break;
          }
          ++len;
          if (!r.advance()) break;
        }
      }
      n_steps += len;
      const uint32_t h = hit ? 1u : 0u;
      const uint32_t enc = (epoch << 26) | (tau << 1) | h;
      const uint32_t senc = ((qc + 2u) << 1) | h;
      if (len == 0) {  // occupied start: the caller's own assignment occludedBy_[idx] = idx
        const size_t L = lin(i, j, k);
        memo[L] = std::max(memo[L], enc);
        snap[qc] = std::max(snap[qc], senc);
      } else {
        VxRay r;
        r.begin(g, i, j, k, e[0], e[1], e[2]);
        for (uint32_t s = 0; s < len; ++s) {
          const size_t L = lin(r.i, r.j, r.k);
          memo[L] = std::max(memo[L], enc);
          const int32_t q = local_of(w, r.i, r.j, r.k);
          if (q >= int32_t(qc)) snap[size_t(q)] = std::max(snap[size_t(q)], senc);
          if (s + 1 < len) r.advance();
        }
      }
    }
    // ---- D
    for (uint32_t q = 0; q < n; ++q) {
      if (skip[q]) continue;
      int32_t i, j, k;
      cell_of(w, q, i, j, k);
      const size_t L = lin(i, j, k);
      if (snap[q] == 0) ++n_unset;
      const bool hit = snap[q] & 1u;
      if (inv) {
        if (!hit) setbit(alive, L, false);
      } else {
        setbit(occl, L, hit);
      }
    }
  }

  void run_pass(const float* e, bool inv) {
    epoch += 1;
    if (epoch >= 64) {
      std::fill(memo.begin(), memo.end(), 0u);
      epoch = 1;
    }
    const int32_t shells = vx_num_shells(g.nx, g.ny);
    const int32_t cols_per_wave = std::max<int32_t>(1, int32_t(wave_cells) / g.nz);
    uint32_t tau = 0;
    for (int32_t o = 0; o < shells; ++o)
      for (int face = 0; face < 3; ++face) {
        const int32_t outer = face == 0 ? g.ny : g.nx - o - 1;
        const int32_t fixed = face == 0 ? g.nx - 1 - o : (face == 1 ? o : g.ny - 1 - o);
        for (int32_t a = 0; a < outer; a += cols_per_wave) {
          Wave w;
          w.face = face;
          w.o = o;
          w.a_begin = a;
          w.a_end = std::min(outer, a + cols_per_wave);
          w.tau_base = tau;
          w.fixed = fixed;
          run_wave(w, e, inv);
          tau += uint32_t(w.a_end - w.a_begin) * uint32_t(g.nz);
        }
      }
  }
};

}  // namespace

extern "C" {

// occ_bytes: one byte per voxel (non-zero = occupied) in OUTPUT order L = (x*ny + y)*nz + z.
// endpoints: n_endpoints x 3.  Outputs (bytes 0/1, output order): occluded = occlusions_ > -1,
// invalid = reference isInvalid after all passes.  stats8: rays, steps, waves, waves needing
// resolution, total rounds, max rounds, prefix steps, reserved.
void wm_run(const int32_t* size3, float res, const float* off3, const uint8_t* occ_bytes, uint32_t n_endpoints,
            const float* endpoints, uint32_t wave_cells, uint64_t seed, uint8_t* occluded, uint8_t* invalid, uint64_t* stats8) {
  Model m;
  m.g.nx = size3[0];
  m.g.ny = size3[1];
  m.g.nz = size3[2];
  m.g.res = res;
  m.g.off[0] = off3[0];
  m.g.off[1] = off3[1];
  m.g.off[2] = off3[2];
  m.nvox = size_t(size3[0]) * size_t(size3[1]) * size_t(size3[2]);
  const size_t words = (m.nvox + 31) / 32;
  m.occ.assign(words, 0);
  m.alive.assign(words, 0);
  m.occl.assign(words, 0);
  m.memo.assign(m.nvox, 0);
  m.wave_cells = wave_cells;
  m.rng = seed * 2 + 1;
  for (size_t L = 0; L < m.nvox; ++L)
    if (occ_bytes[L]) Model::setbit(m.occ, L, true);
  const float origin[3] = {0, 0, 0};
  m.run_pass(origin, false);
  m.alive = m.occl;  // invalid_ = occlusions_
  for (uint32_t p = 0; p < n_endpoints; ++p) m.run_pass(endpoints + 3 * size_t(p), true);
  for (size_t L = 0; L < m.nvox; ++L) {
    occluded[L] = Model::bit(m.occl, L);
    invalid[L] = Model::bit(m.alive, L) && !Model::bit(m.occ, L);
  }
  if (stats8) {
    stats8[0] = m.n_rays;
    stats8[1] = m.n_steps;
    stats8[2] = m.n_waves;
    stats8[3] = m.n_resolve_waves;
    stats8[4] = m.n_rounds;
    stats8[5] = m.max_rounds;
    stats8[6] = m.n_prefix_steps;
    stats8[7] = m.n_unset;
  }
}

}  // extern "C"

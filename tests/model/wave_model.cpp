// TEST INFRASTRUCTURE: sequential CPU model of the wave-parallel visibility algorithm that
// voxelizer_b200/csrc/vx_visibility.cu runs on the GPU ("K4").  It shares the product's
// ray arithmetic header (vx_dda.h) and executes the kernel's phases in an ORDER-SCRAMBLED way (work
// lists are processed in pseudo-random order, merges are max-reductions, the resolve loop sees its
// own updates immediately) so that a test passing here shows the formulation is independent of
// thread scheduling.  tests/test_wave_model.py checks it bit for bit against the oracle (reference
// VoxelGrid::updateOcclusions / updateInvalid, VoxelGrid.cpp:98-316).
//
// The algorithm, per pass (one endpoint) -- identical to the kernel:
//   cells are visited in the reference order (shell / face / column / z).  A cell is LIVE when the
//   visit can change anything: free (occupied cells answer with their own index without a ray and
//   are never crossed by one) and, in an invalid pass, not yet killed.  A WAVE is a run of whole
//   columns of one face holding at most `max_live` live cells.  Waves are processed in order:
//   A. snapshot: owner[r] = id of the latest ray that crossed live cell r so far in this pass
//      (memo, epoch-tagged); cells nothing crossed yet are CANDIDATES.
//   B. resolve which candidates really cast: c casts iff no CASTING candidate with smaller rank
//      crosses c with the part of its ray that lies inside this wave's face (rays are monotone per
//      axis, so a ray never re-enters the face).  Counter formulation: cnt[c] = number of UNDECIDED
//      earlier candidates crossing c.  killed[c] -> does not cast; cnt[c] == 0 -> casts (kills its
//      successors); a candidate that turns out not to cast releases (decrements) its successors.
//      Candidates whose first step is certainly along the face normal have no successors and skip the walk.
//   C. casters get ray ids in visit order (id_base + rank among the wave's casters) and trace:
//      memo[cell] = max(memo[cell], id); wave cells at or after the caster also merge into owner[];
//      hit[id] = the ray ended on an occupied cell.
//   D. apply the visit rule: occlusion pass  occluded[c] = hit[owner] (last visit wins); invalid
//      pass  hit[owner] == 0 kills the cell.
//
// mode 1 (not in the product kernel yet: experiments/r01_rays_in_flight.patch) additionally keeps the rays of an
// INVALID pass in flight across waves: a ray is cast, walked until it has left its wave's face and parked; before
// a wave's snapshot every parked ray is advanced until it is strictly beyond the wave's plane in its direction of
// travel (plus a random number of steps); the visits are logged and applied when every ray has ended.  "The plane
// lies beyond the ray's end voxel" also counts as safe, guarded by an overshoot detector that re-runs the pass
// without that rule.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../voxelizer_b200/csrc/vx_dda.h"

namespace {

struct Model {
  VxGridGeom g;
  size_t nvox;
  std::vector<uint32_t> occ, live, occl;  // bitmaps over L = (i*ny + j)*nz + k
  std::vector<uint32_t> memo;              // (epoch << 25) | ray id
  std::vector<uint8_t> hit;                // per ray id of the current pass
  uint32_t epoch = 0;
  uint32_t max_live = 2048;
  uint32_t max_cols = 512;
  uint64_t rng = 1;
  // rays in flight (mode 1, invalid passes)
  int mode = 0;
  bool in_flight = false, conservative = false, overshoot = false;
  struct Parked {
    VxRay r;
    uint32_t id;
  };
  std::vector<Parked> pool;
  std::vector<std::pair<size_t, uint32_t>> visit_log;  // (cell, owner) of every live visit of the pass
  uint64_t n_parked = 0, n_redo = 0;
  // mode 2 (the product kernel since round 2): a memo mark is only ever READ by the snapshot of a LIVE cell, and a cell that
  // is dead at the start of a pass stays dead for the pass -- so a ray skips the mark of every cell whose (2^s x 2^s x 1)
  // block held no live cell when the pass began (`blive`, a superset of the live set at any later time of the pass)
  std::vector<uint32_t> blive;
  uint32_t bshift = 3;
  uint64_t n_marks = 0, n_marks_skipped = 0;
  size_t block_index(int32_t i, int32_t j, int32_t k) const {
    const size_t bi = size_t(i) >> bshift, bj = size_t(j) >> bshift, nbj = (size_t(g.ny) >> bshift) + 1;
    return (bi * nbj + bj) * size_t(g.nz) + size_t(k);
  }
  void build_blive() {
    const size_t nbi = (size_t(g.nx) >> bshift) + 1, nbj = (size_t(g.ny) >> bshift) + 1;
    blive.assign((nbi * nbj * size_t(g.nz) + 31) / 32, 0u);
    for (int32_t i = 0; i < g.nx; ++i)
      for (int32_t j = 0; j < g.ny; ++j)
        for (int32_t k = 0; k < g.nz; ++k)
          if (bit(live, lin(i, j, k))) setbit(blive, block_index(i, j, k), true);
  }
  bool mark_wanted(int32_t i, int32_t j, int32_t k) {
    ++n_marks;
    if (mode != 2 || bit(blive, block_index(i, j, k))) return true;
    ++n_marks_skipped;
    return false;
  }
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
  template <class T>
  void shuffle(std::vector<T>& v) {
    for (size_t a = v.size(); a > 1; --a) std::swap(v[a - 1], v[next_rand() % a]);
  }

  struct Wave {
    int face;                // 0: i fixed, outer j; 1/2: j fixed, outer i
    int32_t fixed;           // the fixed coordinate
    int32_t a_begin, a_end;  // outer coordinate range
  };

  // wave-local column of (i,j), or -1
  int32_t col_of(const Wave& w, int32_t i, int32_t j) const {
    const int32_t fx = w.face == 0 ? i : j, outer = w.face == 0 ? j : i;
    if (fx != w.fixed || outer < w.a_begin || outer >= w.a_end) return -1;
    return outer - w.a_begin;
  }

  struct LiveCell {
    int32_t i, j, k;
  };

  // true when the first step of the ray from (i,j,k) towards e is certainly along `normal`
  bool normal_first(int32_t i, int32_t j, int32_t k, const float* e, int normal) const {
    const float d[3] = {std::fabs(e[0] - vx_voxel_centre(g.off[0], g.res, i)), std::fabs(e[1] - vx_voxel_centre(g.off[1], g.res, j)),
                        std::fabs(e[2] - vx_voxel_centre(g.off[2], g.res, k))};
    const float other = std::max(d[(normal + 1) % 3], d[(normal + 2) % 3]);
    return d[normal] > 1.0001f * other;
  }

  // one iteration of the trace loop for a ray that is parked on a cell it has not examined yet; true = the ray ended
  bool step_parked(Parked& p) {
    VxRay& r = p.r;
    const size_t L = lin(r.i, r.j, r.k);
    if (bit(occ, L)) {
      hit[p.id] = 1;
      return true;
    }
    memo[L] = std::max(memo[L], (epoch << 25) | p.id);
    ++n_steps;
    if (!r.advance()) return true;
    if (!r.inside(g)) return true;
    if (!conservative) {  // past the end voxel's coordinate?  then "ends before the plane" was not a safe assumption
      if ((r.sx > 0 && r.i > r.ex) || (r.sx < 0 && r.i < r.ex) || (r.sy > 0 && r.j > r.ey) || (r.sy < 0 && r.j < r.ey) ||
          (r.sz > 0 && r.k > r.ez) || (r.sz < 0 && r.k < r.ez))
        overshoot = true;
    }
    return false;
  }
  bool safe_for(const Parked& p, int plane_axis, int32_t c) const {
    const int32_t pa = plane_axis == 0 ? p.r.i : p.r.j, s = plane_axis == 0 ? p.r.sx : p.r.sy, end = plane_axis == 0 ? p.r.ex : p.r.ey;
    const bool beyond = s > 0 ? pa > c : pa < c;
    const bool ends_before = !conservative && (s > 0 ? c > end : c < end);
    return beyond || ends_before;
  }
  // before the snapshot of a wave on (plane_axis, c); drain: until every ray has ended
  void advance_pool(int plane_axis, int32_t c, bool drain) {
    shuffle(pool);
    std::vector<Parked> keep;
    for (Parked& p : pool) {
      bool ended = false;
      uint32_t extra = drain ? 0 : next_rand() % 4;  // the kernel advances a minimum number of steps per wave
      while (!ended && (drain || !safe_for(p, plane_axis, c) || extra > 0)) {
        if (extra) --extra;
        ended = step_parked(p);
      }
      if (!ended) keep.push_back(p);
    }
    pool.swap(keep);
  }

  void run_wave(const Wave& w, const float* e, bool inv, uint32_t& id_base) {
    const int normal = w.face == 0 ? 0 : 1;
    // ---- A: live cells in visit order, rank lookup, snapshot
    std::vector<LiveCell> cells;
    std::vector<int32_t> rank_of(size_t(w.a_end - w.a_begin) * size_t(g.nz), -1);
    for (int32_t a = w.a_begin; a < w.a_end; ++a)
      for (int32_t k = 0; k < g.nz; ++k) {
        const int32_t i = w.face == 0 ? w.fixed : a, j = w.face == 0 ? a : w.fixed;
        if (bit(live, lin(i, j, k))) {
          rank_of[size_t(a - w.a_begin) * size_t(g.nz) + size_t(k)] = int32_t(cells.size());
          cells.push_back({i, j, k});
        }
      }
    const uint32_t n = uint32_t(cells.size());
    ++n_waves;
    if (in_flight) advance_pool(w.face == 0 ? 0 : 1, w.fixed, false);
    const uint32_t SET = 0x80000000u;
    std::vector<uint32_t> owner(n, 0);  // 0 = candidate; SET | id otherwise.  Doubles as cnt[] during B.
    std::vector<uint32_t> cand;
    for (uint32_t r = 0; r < n; ++r) {
      const uint32_t m = memo[lin(cells[r].i, cells[r].j, cells[r].k)];
      if ((m >> 25) == epoch) owner[r] = SET | (m & 0x1FFFFFFu);
      else cand.push_back(r);
    }
    auto live_rank = [&](int32_t i, int32_t j, int32_t k) -> int32_t {
      const int32_t c = col_of(w, i, j);
      return c < 0 ? -1 : rank_of[size_t(c) * size_t(g.nz) + size_t(k)];
    };
    // in-face part of candidate rc's ray: calls f(rank) for live wave cells with larger rank
    auto prefix_walk = [&](uint32_t rc, auto f) {
      VxRay r;
      r.begin(g, cells[rc].i, cells[rc].j, cells[rc].k, e[0], e[1], e[2]);
      for (;;) {
        if (r.next_axis() == normal) return;
        if (!r.advance()) return;
        if (!r.inside(g)) return;
        ++n_prefix_steps;
        if (bit(occ, lin(r.i, r.j, r.k))) return;
        const int32_t x = live_rank(r.i, r.j, r.k);
        if (x > int32_t(rc)) f(uint32_t(x));
      }
    };
    // ---- B: counters
    std::vector<uint32_t> casters;
    if (!cand.empty()) {
      std::vector<uint8_t> status(n, 0), killed(n, 0), has_succ(n, 0);  // status: 0 undecided, 1 cast, 2 not
      bool any = false;
      std::vector<uint32_t> order = cand;
      shuffle(order);
      for (uint32_t rc : order) {
        if (normal_first(cells[rc].i, cells[rc].j, cells[rc].k, e, normal)) continue;
        prefix_walk(rc, [&](uint32_t x) {
          if (!(owner[x] & SET)) {
            owner[x] += 1;
            has_succ[rc] = 1;
            any = true;
          }
        });
      }
      if (any) {
        ++n_resolve_waves;
        uint64_t rounds = 0;
        size_t undecided = cand.size();
        while (undecided) {
          ++rounds;
          shuffle(order);
          for (uint32_t rc : order) {
            if (status[rc]) continue;
            if (killed[rc]) {
              status[rc] = 2;
              --undecided;
              if (has_succ[rc]) prefix_walk(rc, [&](uint32_t x) { if (!(owner[x] & SET)) owner[x] -= 1; });
            } else if (owner[rc] == 0) {
              status[rc] = 1;
              --undecided;
              if (has_succ[rc]) prefix_walk(rc, [&](uint32_t x) { if (!(owner[x] & SET)) killed[x] = 1; });
            }
          }
        }
        n_rounds += rounds;
        max_rounds = std::max(max_rounds, rounds);
        for (uint32_t rc : cand) {
          owner[rc] = 0;
          if (status[rc] == 1) casters.push_back(rc);
        }
      } else {
        casters = cand;
      }
    }
    // ---- C: ids in visit order, traced in scrambled order
    std::vector<uint32_t> slot(casters.size());
    for (uint32_t s = 0; s < slot.size(); ++s) slot[s] = s;
    shuffle(slot);
    if (hit.size() < size_t(id_base) + casters.size()) hit.resize(size_t(id_base) + casters.size() + 1024, 0);
    for (uint32_t s : slot) {
      const uint32_t rc = casters[s];
      const uint32_t id = id_base + s;
      const uint32_t enc = (epoch << 25) | id;
      ++n_rays;
      VxRay r;
      r.begin(g, cells[rc].i, cells[rc].j, cells[rc].k, e[0], e[1], e[2]);
      bool h = false, in_face = true, parked = false;
      for (;;) {
        if (in_flight && !in_face && next_rand() % 3 != 0) {  // park it (sometimes a few steps later)
          pool.push_back(Parked{r, id});
          ++n_parked;
          parked = true;
          break;
        }
        if (!r.inside(g)) break;
        const size_t L = lin(r.i, r.j, r.k);
        if (bit(occ, L)) {
          h = true;
          break;
        }
        if (mark_wanted(r.i, r.j, r.k)) memo[L] = std::max(memo[L], enc);
        if (in_face) {
          const int32_t x = live_rank(r.i, r.j, r.k);
          if (x >= int32_t(rc)) owner[size_t(x)] = std::max(owner[size_t(x)], SET | id);
        }
        ++n_steps;
        if (r.next_axis() == normal) in_face = false;
        if (!r.advance()) break;
      }
      if (!parked) hit[id] = h ? 1 : 0;
    }
    id_base += uint32_t(casters.size());
    // ---- D
    for (uint32_t r = 0; r < n; ++r) {
      const size_t L = lin(cells[r].i, cells[r].j, cells[r].k);
      if (!(owner[r] & SET)) {
        ++n_unset;
        continue;
      }
      if (in_flight) {  // the owner may still be in flight: the pass applies its log at its end
        visit_log.emplace_back(L, owner[r] & 0x1FFFFFFu);
        continue;
      }
      const bool h = hit[owner[r] & 0x1FFFFFFu] != 0;
      if (inv) {
        if (!h) setbit(live, L, false);
      } else {
        setbit(occl, L, h);
      }
    }
  }

  // one face = a sequence of columns, cut into waves of at most max_live live cells / max_cols columns
  void run_face(int face, int32_t fixed, int32_t outer, const float* e, bool inv, uint32_t& id_base) {
    int32_t a = 0;
    while (a < outer) {
      uint32_t lv = 0;
      int32_t b = a;
      while (b < outer && uint32_t(b - a) < max_cols) {
        uint32_t c = 0;
        for (int32_t k = 0; k < g.nz; ++k) {
          const int32_t i = face == 0 ? fixed : b, j = face == 0 ? b : fixed;
          c += bit(live, lin(i, j, k));
        }
        if (b > a && lv + c > max_live) break;
        lv += c;
        ++b;
      }
      Wave w{face, fixed, a, b};
      if (lv) run_wave(w, e, inv, id_base);
      a = b;
    }
  }

  void bump_epoch() {
    epoch += 1;
    if (epoch >= 127) {
      std::fill(memo.begin(), memo.end(), 0u);
      epoch = 1;
    }
  }

  void run_pass(const float* e, bool inv) {
    in_flight = mode == 1 && inv;
    conservative = false;
    const uint64_t rays0 = n_rays, steps0 = n_steps;
    for (;;) {
      bump_epoch();  // (a new epoch also forgets the marks of an abandoned attempt)
      if (mode == 2) build_blive();
      std::fill(hit.begin(), hit.end(), 0);
      pool.clear();
      visit_log.clear();
      overshoot = false;
      uint32_t id_base = 0;
      const int32_t shells = vx_num_shells(g.nx, g.ny);
      for (int32_t o = 0; o < shells; ++o) {
        run_face(0, g.nx - 1 - o, g.ny, e, inv, id_base);
        run_face(1, o, g.nx - o - 1, e, inv, id_base);
        run_face(2, g.ny - 1 - o, g.nx - o - 1, e, inv, id_base);
      }
      if (!in_flight) return;
      advance_pool(0, 0, true);
      if (overshoot && !conservative) {  // nothing of the pass has been applied yet: run it again, conservatively
        conservative = true;
        n_rays = rays0;
        n_steps = steps0;
        ++n_redo;
        continue;
      }
      for (const auto& v : visit_log)
        if (!hit[v.second]) setbit(live, v.first, false);
      return;
    }
  }
};

}  // namespace

extern "C" {

// occ_bytes: one byte per voxel (non-zero = occupied) in OUTPUT order L = (x*ny + y)*nz + z.
// endpoints: n_endpoints x 3.  max_live: live cells per wave (columns are never split).
// Outputs (bytes 0/1, output order): occluded = occlusions_ > -1, invalid = reference isInvalid after
// all passes.  stats8: rays, steps, waves, waves needing resolution, total rounds, max rounds, prefix
// steps, live cells left without an owner (must be 0).
void wm_run_mode(int mode, const int32_t* size3, float res, const float* off3, const uint8_t* occ_bytes, uint32_t n_endpoints,
                 const float* endpoints, uint32_t max_live, uint64_t seed, uint8_t* occluded, uint8_t* invalid, uint64_t* stats8,
                 uint64_t* extra2);

void wm_run(const int32_t* size3, float res, const float* off3, const uint8_t* occ_bytes, uint32_t n_endpoints,
            const float* endpoints, uint32_t max_live, uint64_t seed, uint8_t* occluded, uint8_t* invalid, uint64_t* stats8) {
  wm_run_mode(0, size3, res, off3, occ_bytes, n_endpoints, endpoints, max_live, seed, occluded, invalid, stats8, nullptr);
}

// mode 1: rays of the invalid passes in flight across waves.  extra2: rays parked, passes run again conservatively.
// mode 2: memo marks filtered by the block-level live set of the pass start.  extra2: marks skipped, marks in all.
// The block size is 2^(max_live >> 24) when that is non-zero (the upper byte of max_live; default 8 x 8 x 1).
void wm_run_mode(int mode, const int32_t* size3, float res, const float* off3, const uint8_t* occ_bytes, uint32_t n_endpoints,
                 const float* endpoints, uint32_t max_live, uint64_t seed, uint8_t* occluded, uint8_t* invalid, uint64_t* stats8,
                 uint64_t* extra2) {
  Model m;
  m.mode = mode;
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
  m.live.assign(words, 0);
  m.occl.assign(words, 0);
  m.memo.assign(m.nvox, 0);
  if (max_live >> 24) m.bshift = max_live >> 24;
  max_live &= 0xFFFFFFu;
  m.max_live = max_live < 1 ? 1 : max_live;
  m.rng = seed * 2 + 1;
  for (size_t L = 0; L < m.nvox; ++L)
    if (occ_bytes[L]) Model::setbit(m.occ, L, true);
  // occupied cells: occludedBy returns their own index at once (VoxelGrid.cpp:246-247 via 286-296), so
  // occlusions_ > -1 and invalid_ == own index for good; no ray ever records them.  They are not live.
  for (size_t L = 0; L < m.nvox; ++L) Model::setbit(m.live, L, !Model::bit(m.occ, L));
  m.occl = m.occ;
  const float origin[3] = {0, 0, 0};
  m.run_pass(origin, false);
  for (size_t w = 0; w < words; ++w) m.live[w] = m.occl[w] & ~m.occ[w];  // invalid_ = occlusions_
  for (uint32_t p = 0; p < n_endpoints; ++p) m.run_pass(endpoints + 3 * size_t(p), true);
  for (size_t L = 0; L < m.nvox; ++L) {
    occluded[L] = Model::bit(m.occl, L);
    invalid[L] = Model::bit(m.live, L);
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
  if (extra2) {
    extra2[0] = mode == 2 ? m.n_marks_skipped : m.n_parked;
    extra2[1] = mode == 2 ? m.n_marks : m.n_redo;
  }
}

}  // extern "C"

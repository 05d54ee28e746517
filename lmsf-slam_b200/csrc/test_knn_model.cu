// test_knn_model.cu — the cooperative 5-NN search of knn.cuh executed on the CPU against brute force.
//
// Test infrastructure (run by tests/test_knn_model.py; no GPU needed).  The search is written once as templates over an
// executor; here the same source runs (a) with one lane per query (HostGroup1) and (b) as a full emulated warp:
// 32 host threads in lockstep, four groups of eight lanes, every ballot / shuffle / any a barrier-synchronised
// exchange — the lane arithmetic (prefix sums, ballots, rank counting, list compression) is the code the device runs.
// The index is built on the CPU in the layout mapindex.cu produces (hash table of 1 m cells, L1 records, L2 starts,
// points grouped by (L0, L1, L2) with x fastest).
//
//   ./test_knn_model [n_map] [n_query] [seed]     exit code 0 = every query identical to brute force (ids and
//                                                 distance bits), in both executors, unseeded and seeded
#include <algorithm>
#include <barrier>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <thread>

#include "knn.cuh"

using namespace lm;

// ---------------------------------------------------------------- CPU build of the index
struct HostIndex {
  std::vector<float4> sorted;
  std::vector<CellRec> table;
  std::vector<L1Rec> l1;
  std::vector<int> l2_start;
  MapDev md;
  MapView view() const {
    MapView v;
    v.sorted = sorted.data();
    v.table = table.data();
    v.l1 = l1.data();
    v.l2_start = l2_start.data();
    v.dev = &md;
    return v;
  }
};

static unsigned sub_cell_host(const float4& p) {
  int ax = (int)floorf(p.x * 16.0f), ay = (int)floorf(p.y * 16.0f), az = (int)floorf(p.z * 16.0f);
  unsigned f1 = (((az >> 2) & 3) << 4) | (((ay >> 2) & 3) << 2) | ((ax >> 2) & 3);
  unsigned f2 = ((az & 3) << 4) | ((ay & 3) << 2) | (ax & 3);
  return (f1 << 6) | f2;
}

static void build_index(const std::vector<float4>& pts, HostIndex& ix) {
  MapDev& md = ix.md;
  memset(&md, 0, sizeof md);
  md.n = (int)pts.size();
  float mn[3] = {1e30f, 1e30f, 1e30f}, mx[3] = {-1e30f, -1e30f, -1e30f};
  for (const float4& p : pts) {
    const float v[3] = {p.x, p.y, p.z};
    for (int a = 0; a < 3; ++a) {
      mn[a] = std::min(mn[a], v[a]);
      mx[a] = std::max(mx[a], v[a]);
    }
  }
  for (int a = 0; a < 3; ++a) {
    md.min_c[a] = (int)floorf(mn[a]);
    md.dim[a] = (int)floorf(mx[a]) - md.min_c[a] + 1;
    int b = 0;
    while ((1LL << b) < (long long)md.dim[a]) ++b;
    md.bits[a] = b;
  }
  // points per L0 cell
  std::map<unsigned long long, std::vector<int>> cells;
  for (int i = 0; i < (int)pts.size(); ++i) {
    const float4& p = pts[i];
    int ax = (int)floorf(p.x * 16.0f), ay = (int)floorf(p.y * 16.0f), az = (int)floorf(p.z * 16.0f);
    cells[pack_cell(md, (ax >> 4) - md.min_c[0], (ay >> 4) - md.min_c[1], (az >> 4) - md.min_c[2])].push_back(i);
  }
  unsigned slots = 64;
  while (slots < 2 * cells.size()) slots <<= 1;
  md.table_mask = slots - 1;
  CellRec empty;
  memset(&empty, 0xff, sizeof empty);
  ix.table.assign(slots, empty);
  ix.sorted.clear();
  ix.l1.clear();
  ix.l2_start.clear();
  for (auto& kv : cells) {
    unsigned h = hash_cell(kv.first) & md.table_mask;
    while (ix.table[h].key != ~0ull) h = (h + 1) & md.table_mask;
    CellRec& r = ix.table[h];
    r.key = kv.first;
    r.start = (int)ix.sorted.size();
    r.fine_base = (int)ix.l1.size();
    r.mask = 0ull;
    r.pad = 0;
    std::vector<int>& ids = kv.second;
    std::stable_sort(ids.begin(), ids.end(), [&](int a, int b) { return sub_cell_host(pts[a]) < sub_cell_host(pts[b]); });
    int prev = -1;
    for (int id : ids) {
      const int sc = (int)sub_cell_host(pts[id]);
      if (sc != prev) {
        const int f1 = sc >> 6, f2 = sc & 63;
        if (prev < 0 || (prev >> 6) != f1) {
          L1Rec lr;
          lr.mask = 0ull;
          lr.first = (int)ix.l2_start.size();
          lr.pad = 0;
          ix.l1.push_back(lr);
          r.mask |= 1ull << f1;
        }
        ix.l1.back().mask |= 1ull << f2;
        ix.l2_start.push_back((int)ix.sorted.size());
        prev = sc;
      }
      float4 p = pts[id];
      memcpy(&p.w, &id, 4);
      ix.sorted.push_back(p);
    }
    r.end = (int)ix.sorted.size();
    ix.l2_start.push_back(r.end);  // sentinel of this L0 cell
  }
  md.n_fine = (int)ix.l1.size();
}

// ---------------------------------------------------------------- brute force (FLANN's L2_Simple, ties by index)
static int brute5(const std::vector<float4>& pts, float qx, float qy, float qz, unsigned long long out[5]) {
  std::vector<unsigned long long> keys;
  for (int i = 0; i < (int)pts.size(); ++i) {
    const float dx = pts[i].x - qx, dy = pts[i].y - qy, dz = pts[i].z - qz;
    float r = dx * dx;
    r = r + dy * dy;
    r = r + dz * dz;
    if (r < 1.0f) keys.push_back(kg_key(r, i));
  }
  const int n = (int)std::min<size_t>(5, keys.size());
  std::partial_sort(keys.begin(), keys.begin() + n, keys.end());
  for (int k = 0; k < n; ++k) out[k] = keys[k];
  return n;
}

// ---------------------------------------------------------------- emulated warp
struct EmuWarp {
  std::barrier<> bar{32};
  unsigned long long buf[2][32];
};
struct EmuGroup8 {
  static constexpr int G = 8;
  int l;
  unsigned gshift;
  int lane;
  EmuWarp* w;
  mutable int phase = 0;
  EmuGroup8(EmuWarp* w_, int lane_) : l(lane_ & 7), gshift(lane_ & 24), lane(lane_), w(w_) {}
  // every lane deposits a value, all wait, every lane may read any
  HD const unsigned long long* exchange(unsigned long long v) const {
#ifndef __CUDA_ARCH__
    unsigned long long* b = w->buf[phase];
    phase ^= 1;
    b[lane] = v;
    w->bar.arrive_and_wait();
    return b;
#else
    return nullptr;
#endif
  }
  HD unsigned ballot(bool p) const {
    const unsigned long long* b = exchange(p ? 1ull : 0ull);
    unsigned m = 0;
    for (int i = 0; i < 8; ++i) m |= (unsigned)b[gshift + i] << i;
    return m;
  }
  HD bool any(bool p) const {
    const unsigned long long* b = exchange(p ? 1ull : 0ull);
    for (int i = 0; i < 32; ++i)
      if (b[i]) return true;
    return false;
  }
  HD int shfl(int v, int src) const { return (int)(unsigned)exchange((unsigned)v)[gshift + (src & 7)]; }
  HD int shfl_up(int v, int d) const {
    const unsigned long long* b = exchange((unsigned)v);
    return l >= d ? (int)(unsigned)b[lane - d] : v;
  }
  HD unsigned long long shfl_xor64(unsigned long long v, int d) const { return exchange(v)[gshift + ((l ^ d) & 7)]; }
  HD void sync() const { exchange(0ull); }
};

struct Query {
  float x, y, z;
  unsigned long long seed;
};

static long run_single(const HostIndex& ix, const std::vector<float4>& pts, const std::vector<Query>& qs, bool seeded) {
  long bad = 0;
  const MapView mv = ix.view();
  KnnScratch s;
  HostGroup1 x;
  for (const Query& q : qs) {
    unsigned long long ref[5];
    const int nr = brute5(pts, q.x, q.y, q.z, ref);
    const int n = kg_knn5(x, mv, &s, q.x, q.y, q.z, true, seeded ? q.seed : 0ull);
    bool ok = n == nr;
    for (int k = 0; ok && k < n; ++k) ok = s.acc[k] == ref[k];
    if (!ok && bad++ < 5) fprintf(stderr, "single%s: query (%g %g %g): n %d vs %d\n", seeded ? " seeded" : "", q.x, q.y, q.z, n, nr);
  }
  return bad;
}

static long run_warp(const HostIndex& ix, const std::vector<float4>& pts, const std::vector<Query>& qs, bool seeded) {
  const MapView mv = ix.view();
  EmuWarp w;
  static KnnScratch scratch[4];
  std::vector<int> n_out(qs.size(), -1);
  std::vector<std::array<unsigned long long, 5>> res(qs.size());
  auto body = [&](int lane) {
    EmuGroup8 x(&w, lane);
    const int g = lane >> 3;
    for (size_t base = 0; base < qs.size(); base += 4) {
      const size_t qi = base + g;
      const bool active = qi < qs.size();
      const Query q = active ? qs[qi] : Query{0, 0, 0, 0};
      const int n = kg_knn5(x, mv, &scratch[g], q.x, q.y, q.z, active, seeded ? q.seed : 0ull);
      if (active && x.l == 0) {
        n_out[qi] = n;
        for (int k = 0; k < n; ++k) res[qi][k] = scratch[g].acc[k];
      }
      x.sync();
    }
  };
  std::vector<std::thread> th;
  for (int lane = 0; lane < 32; ++lane) th.emplace_back(body, lane);
  for (auto& t : th) t.join();
  long bad = 0;
  for (size_t i = 0; i < qs.size(); ++i) {
    unsigned long long ref[5];
    const int nr = brute5(pts, qs[i].x, qs[i].y, qs[i].z, ref);
    bool ok = n_out[i] == nr;
    for (int k = 0; ok && k < nr; ++k) ok = res[i][k] == ref[k];
    if (!ok && bad++ < 5)
      fprintf(stderr, "warp%s: query %zu (%g %g %g): n %d vs %d\n", seeded ? " seeded" : "", i, qs[i].x, qs[i].y, qs[i].z, n_out[i], nr);
  }
  return bad;
}

int main(int argc, char** argv) {
  const int n_map = argc > 1 ? atoi(argv[1]) : 60000;
  const int n_q = argc > 2 ? atoi(argv[2]) : 1500;
  const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
  std::mt19937 rng(seed);
  std::uniform_real_distribution<float> U(0.f, 1.f);
  std::normal_distribution<float> N(0.f, 1.f);
  // a scene with every density regime: a dense ground patch, a very dense wall, a thin pole, sparse clutter,
  // exact duplicates (distance ties) and points on cell boundaries
  std::vector<float4> pts;
  auto add = [&](float x, float y, float z) { pts.push_back(make_float4(x, y, z, 0.f)); };
  const int n_ground = n_map * 4 / 10, n_wall = n_map * 3 / 10, n_pole = n_map / 20, n_dup = n_map / 20;
  for (int i = 0; i < n_ground; ++i) add(-12.f + 24.f * U(rng), -12.f + 24.f * U(rng), -1.7f + 0.01f * N(rng));
  for (int i = 0; i < n_wall; ++i) add(5.0f + 0.005f * N(rng), -3.f + 6.f * U(rng), -1.7f + 3.f * U(rng));
  for (int i = 0; i < n_pole; ++i) add(-4.f + 0.02f * N(rng), 3.f + 0.02f * N(rng), -1.7f + 4.f * U(rng));
  for (int i = 0; i < n_dup; ++i) {
    const float4 p = pts[(size_t)(U(rng) * (float)(pts.size() - 1))];
    add(p.x, p.y, p.z);
  }
  for (int i = 0; i < n_map / 40; ++i)  // lattice points: on cell faces at every level
    add(0.0625f * (float)(int)(-64.f + 128.f * U(rng)), 0.25f * (float)(int)(-16.f + 32.f * U(rng)), (float)(int)(-2.f + 4.f * U(rng)));
  while ((int)pts.size() < n_map) add(-40.f + 80.f * U(rng), -40.f + 80.f * U(rng), -3.f + 10.f * U(rng));
  HostIndex ix;
  build_index(pts, ix);

  std::vector<Query> qs;
  for (int i = 0; i < n_q; ++i) {
    Query q;
    const float pick = U(rng);
    if (pick < 0.6f) {  // near a map point (what registration asks)
      const float4 p = pts[(size_t)(U(rng) * (float)(pts.size() - 1))];
      const float s = pick < 0.3f ? 0.02f : 0.3f;
      q.x = p.x + s * N(rng), q.y = p.y + s * N(rng), q.z = p.z + s * N(rng);
    } else if (pick < 0.7f) {  // exactly a map point / a lattice point
      const float4 p = pts[(size_t)(U(rng) * (float)(pts.size() - 1))];
      q.x = p.x, q.y = p.y, q.z = p.z;
    } else if (pick < 0.95f) {  // anywhere in the scene, mostly sparse
      q.x = -42.f + 84.f * U(rng), q.y = -42.f + 84.f * U(rng), q.z = -4.f + 12.f * U(rng);
    } else {  // far outside the grid
      q.x = 300.f * N(rng), q.y = 300.f * N(rng), q.z = 100.f * N(rng);
    }
    q.seed = 0ull;
    qs.push_back(q);
  }
  // seeds as the registration provides them: the five neighbours of a slightly moved query -> an inclusive bound
  // (the largest of their keys at the new position); some seeds beyond the search radius, some queries unseeded
  for (Query& q : qs) {
    unsigned long long nb[5];
    const float mx = q.x + 0.03f * N(rng), my = q.y + 0.03f * N(rng), mz = q.z + 0.03f * N(rng);
    if (brute5(pts, mx, my, mz, nb) == 5 && U(rng) < 0.9f) {
      unsigned long long k5 = 0ull;
      for (int k = 0; k < 5; ++k) {
        const float4 p = pts[kg_key_id(nb[k])];
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        float r = dx * dx;
        r = r + dy * dy;
        r = r + dz * dz;
        k5 = std::max(k5, kg_key(r, kg_key_id(nb[k])));
      }
      q.seed = k5;
    }
  }
  long bad = 0;
  bad += run_single(ix, pts, qs, false);
  bad += run_single(ix, pts, qs, true);
  bad += run_warp(ix, pts, qs, false);
  bad += run_warp(ix, pts, qs, true);
  printf("map %zu points, %d queries x 4 runs: %ld mismatches\n", pts.size(), n_q, bad);
  return bad ? 1 : 0;
}

// test_knn_model.cu — the 5-NN search of knn.cuh executed on the CPU against brute force.
//
// Test infrastructure (run by tests/test_knn_model.py; no GPU needed).  The search is plain per-thread code (HD): the
// source the device runs is compiled for the host here, with the per-lane segment list laid out as in shared memory.
// The index is built on the CPU in the layout mapindex.cu produces (hash table of 1 m cells, L1 records, L2 starts,
// points grouped by (L0, L1, L2) with x fastest).
//
//   ./test_knn_model [n_map] [n_query] [seed]     exit code 0 = every query identical to brute force (ids and
//                                                 distance bits), unseeded and seeded
#include <algorithm>
#include <barrier>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>

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

struct Query {
  float x, y, z;
  bool seeded;
  unsigned long long seed[5];
};

// ---------------------------------------------------------------- emulated warp: 32 host threads in lockstep, every
// ballot / shuffle a barrier-synchronised exchange (kw_knn5 is written over this interface)
struct EmuWarp {
  std::barrier<> bar{32};
  unsigned long long buf[2][32];
};
struct Emu32 {
  int lane;
  EmuWarp* w;
  mutable int phase = 0;
  Emu32(EmuWarp* w_, int lane_) : lane(lane_), w(w_) {}
  const unsigned long long* exchange(unsigned long long v) const {
    unsigned long long* b = w->buf[phase];
    phase ^= 1;
    b[lane] = v;
    w->bar.arrive_and_wait();
    return b;
  }
  unsigned ballot(bool p) const {
    const unsigned long long* b = exchange(p ? 1ull : 0ull);
    unsigned m = 0;
    for (int i = 0; i < 32; ++i) m |= (unsigned)b[i] << i;
    return m;
  }
  int shfl(int v, int src) const { return (int)(unsigned)exchange((unsigned)v)[src & 31]; }
  int shfl_up(int v, int d) const {
    const unsigned long long* b = exchange((unsigned)v);
    return lane >= d ? (int)(unsigned)b[lane - d] : v;
  }
  unsigned long long shfl64(unsigned long long v, int src) const { return exchange(v)[src & 31]; }
  unsigned long long shfl_xor64(unsigned long long v, int d) const { return exchange(v)[(lane ^ d) & 31]; }
  void sync() const { exchange(0ull); }
};

// the registration kernel's pipeline: the per-thread search with deferral, the deferred queries by the (emulated) warp
static long run_deferred(const HostIndex& ix, const std::vector<float4>& pts, const std::vector<Query>& qs, long* n_def) {
  const MapView mv = ix.view();
  std::vector<int> smem(32 * 2 * KQ_SEG_CAP, -1);
  KqList li;
  li.seg = smem.data();
  li.stride = 32;
  std::vector<int> deferred;
  long bad = 0;
  auto check = [&](size_t i, int n, const KqTop& top, const char* what) {
    unsigned long long ref[5];
    const int nr = brute5(pts, qs[i].x, qs[i].y, qs[i].z, ref);
    bool ok = n == nr;
    for (int k = 0; ok && k < n; ++k) ok = top.k[k] == ref[k];
    if (!ok && bad++ < 5) fprintf(stderr, "%s: query %zu (%g %g %g): n %d vs %d\n", what, i, qs[i].x, qs[i].y, qs[i].z, n, nr);
  };
  for (size_t i = 0; i < qs.size(); ++i) {
    KqTop top;
    const int n = kq_knn5<true>(mv, li, qs[i].x, qs[i].y, qs[i].z, nullptr, top);
    if (n < 0)
      deferred.push_back((int)i);
    else
      check(i, n, top, "thread");
  }
  *n_def = (long)deferred.size();
  EmuWarp w;
  static KwScratch scratch;
  std::vector<KqTop> tops(deferred.size());
  std::vector<int> ns(deferred.size(), -2);
  auto body = [&](int lane) {
    Emu32 x(&w, lane);
    for (size_t d = 0; d < deferred.size(); ++d) {
      const Query& q = qs[deferred[d]];
      KqTop top;
      const int n = kw_knn5(x, mv, &scratch, q.x, q.y, q.z, top);
      if (lane == (int)(d % 32)) tops[d] = top, ns[d] = n;  // every lane holds the same list: take a different one each time
      x.sync();
    }
  };
  std::vector<std::thread> th;
  for (int lane = 0; lane < 32; ++lane) th.emplace_back(body, lane);
  for (auto& t : th) t.join();
  for (size_t d = 0; d < deferred.size(); ++d) {
    if (ns[d] == -1) {  // item list overflow: the kernel falls back to the complete per-thread search
      KqTop top;
      const int n = kq_knn5<false>(mv, li, qs[deferred[d]].x, qs[deferred[d]].y, qs[deferred[d]].z, nullptr, top);
      check(deferred[d], n, top, "fallback");
    } else {
      check(deferred[d], ns[d], tops[d], "warp");
    }
  }
  return bad;
}

// lane `lane` of a 32-lane list (the device layout: stride 32), to exercise the indexing
static long run(const HostIndex& ix, const std::vector<float4>& pts, const std::vector<Query>& qs, bool seeded) {
  long bad = 0;
  const MapView mv = ix.view();
  std::vector<int> smem(32 * 2 * KQ_SEG_CAP, -1);
  int lane = 0;
  for (const Query& q : qs) {
    KqList li;
    li.seg = smem.data() + lane;
    li.stride = 32;
    lane = (lane + 7) & 31;
    unsigned long long ref[5];
    const int nr = brute5(pts, q.x, q.y, q.z, ref);
    KqTop top;
    const int n = kq_knn5<false>(mv, li, q.x, q.y, q.z, (seeded && q.seeded) ? q.seed : nullptr, top);
    bool ok = n == nr;
    for (int k = 0; ok && k < n; ++k) ok = top.k[k] == ref[k];
    if (!ok && bad++ < 5)
      fprintf(stderr, "%s: query (%g %g %g): n %d vs %d\n", seeded ? "seeded" : "unseeded", q.x, q.y, q.z, n, nr);
  }
  return bad;
}

// work profile on a recorded scene: file = int32 n_map, int32 n_q, n_map x float4, n_q x float3 (tests/make_knn_scene.py)
#ifdef LMSF_KNN_STATS
static int profile_file(const char* path) {
  FILE* f = fopen(path, "rb");
  if (!f) return 2;
  int n_map = 0, n_q = 0;
  if (fread(&n_map, 4, 1, f) != 1 || fread(&n_q, 4, 1, f) != 1) return 2;
  std::vector<float4> pts(n_map);
  std::vector<float> qf(3 * (size_t)n_q);
  if (fread(pts.data(), 16, n_map, f) != (size_t)n_map || fread(qf.data(), 12, n_q, f) != (size_t)n_q) return 2;
  fclose(f);
  HostIndex ix;
  build_index(pts, ix);
  const MapView mv = ix.view();
  std::vector<int> smem(32 * 2 * KQ_SEG_CAP, -1);
  KqList li;
  li.seg = smem.data();
  li.stride = 32;
  std::vector<std::array<unsigned long long, 5>> k5(n_q);
  for (auto& a : k5) a.fill(0ull);
  for (int pass = 0; pass < 2; ++pass) {
    memset(g_knn_stat_host, 0, sizeof g_knn_stat_host);
    std::vector<unsigned long long> cand(n_q), cost(n_q);
    long full = 0;
    for (int i = 0; i < n_q; ++i) {
      const unsigned long long c0 = g_knn_stat_host[2];
      unsigned long long g0[16];
      memcpy(g0, g_knn_stat_host, sizeof g0);
      KqTop top;
      // pass 1 = seeded with the first pass's own 5th key (what the second outer iteration sees, the pose barely moved)
      int n = kq_knn5<true>(mv, li, qf[3 * i], qf[3 * i + 1], qf[3 * i + 2], (pass && k5[i][4] != 0ull) ? k5[i].data() : nullptr, top);
      if (n < 0) {
        ++g_knn_stat_host[15];
        KqTop t2;
        const unsigned long long keep[16] = {g_knn_stat_host[0], g_knn_stat_host[1], g_knn_stat_host[2], g_knn_stat_host[3]};
        n = kq_knn5<false>(mv, li, qf[3 * i], qf[3 * i + 1], qf[3 * i + 2], nullptr, t2);
        top = t2;
        for (int k = 0; k < 4; ++k) g_knn_stat_host[k] = keep[k];
        for (int k = 8; k < 15; ++k) g_knn_stat_host[k] = g0[k];
      }
      if (n == 5) {
        for (int k = 0; k < 5; ++k) k5[i][k] = top.k[k];
        ++full;
      }
      cand[i] = g_knn_stat_host[2] - c0;
      // serial cost model of one query (dependent-latency units): an L0 probe 1, an L1 cell 2, a sweep 2
      unsigned long long sweeps = 0;
      for (int k = 8; k < 16; ++k) sweeps += g_knn_stat_host[k] - g0[k];
      cost[i] = (g_knn_stat_host[0] - g0[0]) + 2 * (g_knn_stat_host[1] - g0[1]) + 2 * sweeps + cand[i] / 8;
      if (cost[i] > 150 && getenv("KNN_PROF_VERBOSE"))
        printf("    query %d (%.2f %.2f %.2f): cost %llu = L0 %llu, L1 %llu, sweeps %llu, cand %llu, n %d d5 %.3f\n", i, qf[3 * i],
               qf[3 * i + 1], qf[3 * i + 2], cost[i], g_knn_stat_host[0] - g0[0], g_knn_stat_host[1] - g0[1], sweeps, cand[i], n,
               n == 5 ? sqrtf(kg_key_d2(top.k[4])) : -1.f);
    }
    {
      std::vector<unsigned long long> cs(cost);
      std::sort(cs.begin(), cs.end());
      unsigned long long tot = 0;
      for (auto v : cs) tot += v;
      printf("  latency-chain model per query: mean %.1f p50 %llu p90 %llu p99 %llu p99.9 %llu max %llu\n", (double)tot / n_q,
             cs[n_q / 2], cs[n_q * 9 / 10], cs[n_q * 99 / 100], cs[(size_t)(n_q * 0.999)], cs[n_q - 1]);
    }
    // per-warp cost of the flat candidate loop: the longest lane of every 32 consecutive queries
    unsigned long long sum = 0, warp_max_sum = 0;
    for (int i = 0; i < n_q; i += 32) {
      unsigned long long mx = 0;
      for (int j = i; j < std::min(n_q, i + 32); ++j) sum += cand[j], mx = std::max(mx, cand[j]);
      warp_max_sum += mx * 32;
    }
    std::sort(cand.begin(), cand.end());
    const unsigned long long* g = g_knn_stat_host;
    printf("%s: %d queries, %ld full | per query: L0 lookups %.2f  L1 cells %.2f  segments %.2f  candidates %.1f "
           "(p50 %llu p90 %llu p99 %llu max %llu) | lane efficiency of the scan %.2f | unseeded %llu seeded %llu "
           "ball %llu sparse %llu deferred %llu\n",
           pass ? "seeded" : "unseeded", n_q, full, (double)g[0] / n_q, (double)g[1] / n_q, (double)g[3] / n_q,
           (double)g[2] / n_q, cand[n_q / 2], cand[n_q * 9 / 10], cand[n_q * 99 / 100], cand[n_q - 1],
           (double)sum / (double)warp_max_sum, g[8], g[9], g[10], g[11], g[15]);
  }
  return 0;
}
#endif

int main(int argc, char** argv) {
#ifdef LMSF_KNN_STATS
  if (argc > 2 && !strcmp(argv[1], "--file")) return profile_file(argv[2]);
#endif
  const int n_map = argc > 1 ? atoi(argv[1]) : 60000;
  const int n_q = argc > 2 ? atoi(argv[2]) : 1500;
  const unsigned seed = argc > 3 ? (unsigned)atoi(argv[3]) : 1u;
  const int grow0 = argc > 4 ? atoi(argv[4]) : 0;  // MapDev::grow0 (the density hint of voxel-filtered maps)
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
  ix.md.grow0 = grow0;

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
    q.seeded = false;
    qs.push_back(q);
  }
  // seeds as the registration provides them: the five neighbours of a slightly moved query -> an inclusive bound
  // (the largest of their keys at the new position); some seeds beyond the search radius, some queries unseeded
  for (Query& q : qs) {
    unsigned long long nb[5];
    const float mx = q.x + 0.03f * N(rng), my = q.y + 0.03f * N(rng), mz = q.z + 0.03f * N(rng);
    if (brute5(pts, mx, my, mz, nb) == 5 && U(rng) < 0.9f) {
      q.seeded = true;
      for (int k = 0; k < 5; ++k) {
        const float4 p = pts[kg_key_id(nb[k])];
        const float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
        float r = dx * dx;
        r = r + dy * dy;
        r = r + dz * dz;
        q.seed[k] = kg_key(r, kg_key_id(nb[k]));
      }
    }
  }
  long bad = 0;
  bad += run(ix, pts, qs, false);
  bad += run(ix, pts, qs, true);
  long n_def = 0;
  bad += run_deferred(ix, pts, qs, &n_def);
  printf("map %zu points, %d queries x 3 runs (%ld searched by the warp): %ld mismatches\n", pts.size(), n_q, n_def, bad);
  return bad ? 1 : 0;
}

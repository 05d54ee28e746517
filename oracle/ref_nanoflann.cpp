// ref_nanoflann.cpp — thin C wrapper that instantiates the REFERENCE'S OWN vendored
// nanoflann 1.3.2 (include/Algorithm/PointClouds/processing/GlobalDescriptor/scanContext/
// nanoflann.hpp, compiled where it lies under /root/reference; never copied) as a 3-D
// exact kd-tree with the L2_Simple metric (nanoflann.hpp:423-446, same fp32 accumulation
// order as FLANN's L2_Simple used by pcl::KdTreeFLANN) and leaf size 15.
//
// TEST INFRASTRUCTURE ONLY: output goes to oracle/_ref/ and is used by tests/ to
// cross-check the oracle's kNN.  Built by oracle/Makefile when /root/reference exists.
#include <cstddef>
#include <cstdint>
#include <limits>
#include <vector>

#include "nanoflann.hpp"
#include "KDTreeVectorOfVectorsAdaptor.h"

namespace {
struct Cloud {
  const float* p;
  size_t n;
  inline size_t kdtree_get_point_count() const { return n; }
  inline float kdtree_get_pt(const size_t idx, const size_t dim) const { return p[idx * 4 + dim]; }
  template <class BBOX>
  bool kdtree_get_bbox(BBOX&) const { return false; }
};
typedef nanoflann::KDTreeSingleIndexAdaptor<nanoflann::L2_Simple_Adaptor<float, Cloud>, Cloud, 3> Tree;
}  // namespace

extern "C" int ref_nanoflann_knn5(const float* xyzi, int n, const float* q_xyz, int nq, int32_t* idx5, float* d2_5) {
  Cloud cloud{xyzi, (size_t)n};
  Tree tree(3, cloud, nanoflann::KDTreeSingleIndexAdaptorParams(15));
  tree.buildIndex();
  for (int i = 0; i < nq; ++i) {
    size_t idx[5];
    float d2[5];
    nanoflann::KNNResultSet<float> rs(5);
    rs.init(idx, d2);
    tree.findNeighbors(rs, q_xyz + 3 * i, nanoflann::SearchParams(10));
    size_t found = rs.size();
    for (size_t k = 0; k < 5; ++k) {
      idx5[5 * i + k] = k < found ? (int32_t)idx[k] : -1;
      d2_5[5 * i + k] = k < found ? d2[k] : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// The reference's own ring-key tree: KDTreeVectorOfVectorsAdaptor<vector<vector<float>>, float> with leaf 10 and
// the default metric_L2 (SceneRecognitionScanContext.hpp:86-91), searched exactly like descFindSimilar (:272-279).
extern "C" int ref_nanoflann_ringkey_knn10(const float* keys, int n, const float* q_keys, int nq, int32_t* idx10,
                                           float* d10) {
  typedef std::vector<std::vector<float> > KeyMat;
  typedef KDTreeVectorOfVectorsAdaptor<KeyMat, float> KeyTree;
  KeyMat mat((size_t)n, std::vector<float>(20));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 20; ++d) mat[i][d] = keys[20 * (size_t)i + d];
  KeyTree tree(20, mat, 10);
  for (int q = 0; q < nq; ++q) {
    std::vector<size_t> idx(10);
    std::vector<float> dist(10);
    nanoflann::KNNResultSet<float> rs(10);
    rs.init(&idx[0], &dist[0]);
    tree.index->findNeighbors(rs, q_keys + 20 * (size_t)q, nanoflann::SearchParams(10));
    for (size_t k = 0; k < 10; ++k) {
      idx10[10 * q + k] = k < rs.size() ? (int32_t)idx[k] : -1;
      d10[10 * q + k] = k < rs.size() ? dist[k] : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// The same tree kept alive between calls (bench.py's loop-closure cpu_baseline: the reference rebuilds its tree every
// tenth keyframe and searches it for every keyframe in between, SceneRecognitionScanContext.hpp:74-92 — the build is
// therefore outside the timed region, the searches inside).
namespace {
struct KeyTreeHolder {
  typedef std::vector<std::vector<float> > KeyMat;
  typedef KDTreeVectorOfVectorsAdaptor<KeyMat, float> KeyTree;
  KeyMat mat;
  KeyTree* tree = nullptr;
  ~KeyTreeHolder() { delete tree; }
};
}  // namespace

extern "C" void* ref_ringkey_tree_create(const float* keys, int n) {
  KeyTreeHolder* h = new KeyTreeHolder();
  h->mat.assign((size_t)n, std::vector<float>(20));
  for (int i = 0; i < n; ++i)
    for (int d = 0; d < 20; ++d) h->mat[i][d] = keys[20 * (size_t)i + d];
  h->tree = new KeyTreeHolder::KeyTree(20, h->mat, 10);
  return h;
}

extern "C" void ref_ringkey_tree_destroy(void* p) { delete static_cast<KeyTreeHolder*>(p); }

extern "C" int ref_ringkey_tree_knn10(void* p, const float* q_keys, int nq, int32_t* idx10, float* d10) {
  KeyTreeHolder* h = static_cast<KeyTreeHolder*>(p);
  for (int q = 0; q < nq; ++q) {
    std::vector<size_t> idx(10);
    std::vector<float> dist(10);
    nanoflann::KNNResultSet<float> rs(10);
    rs.init(&idx[0], &dist[0]);
    h->tree->index->findNeighbors(rs, q_keys + 20 * (size_t)q, nanoflann::SearchParams(10));
    for (size_t k = 0; k < 10; ++k) {
      idx10[10 * q + k] = k < rs.size() ? (int32_t)idx[k] : -1;
      d10[10 * q + k] = k < rs.size() ? dist[k] : std::numeric_limits<float>::infinity();
    }
  }
  return 0;
}

// Frame::ComputeBoW (reference src/Frame.cc:828-833) = DBoW2 TemplatedVocabulary::transform(features, BowVector&,
// FeatureVector&, levelsup) (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1126-1194) on the descriptors an ORB handle's
// last batch left on the device.  Two kernels:
//   k_bow_descend  one thread per descriptor walks the vocabulary tree (:1217-1258).  The children of a node are stored
//                  next to each other (k * 32 contiguous bytes per step), the whole tree (35 MB for the 10^6-word ORB
//                  vocabulary) stays in the 126 MB L2 across frames.
//   k_bow_reduce   one CTA per frame turns the (word, node) pairs into the two std::maps: a bitonic sort of
//                  (id << 32 | descriptor index) in shared memory, run heads = the map's keys in iteration order,
//                  BowVector::addWeight's repeated addition per run (BowVector.cpp:34-46), BowVector::normalize's
//                  sum in key order by one thread (:62-84; the order is part of the result), division by all.
#include <algorithm>
#include <cmath>
#include <mutex>
#include <vector>

#include "drfe_internal.h"

namespace drfe {

struct VocabDev {
  const int* child_start;     // [nnodes + 2] into the child arrays, by node id (0 = root)
  const int* child_id;        // [nnodes]     node id of the child in slot j
  const uint4* child_desc;    // [nnodes][2]  its descriptor
  const int* word_of;         // [nnodes + 1] m_nodes[id].word_id
  const double* weight;       // [nnodes + 1] m_nodes[id].weight
  int nid_level;              // m_L - levelsup
};

__global__ void __launch_bounds__(128) k_bow_descend(VocabDev V, const uint8_t* __restrict__ desc, const int* __restrict__ cnt, int cap,
                                                     int* __restrict__ word_id, int* __restrict__ node_id, double* __restrict__ wgt) {
  const int f = blockIdx.y, i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= min(cnt[f], cap)) return;
  const long long o = (long long)f * cap + i;
  const uint4* dp = reinterpret_cast<const uint4*>(desc + o * 32);
  const uint4 a = dp[0], b = dp[1];
  int node = 0, level = 0, nid = 0;
  bool nid_set = V.nid_level <= 0;                                 // :1227 root
  int j0 = V.child_start[0], j1 = V.child_start[1];
  do {                                                             // :1232-1253
    ++level;
    int best = 1 << 30, best_j = j0;
    for (int j = j0; j < j1; ++j) {
      const uint4 c = V.child_desc[2 * j], d = V.child_desc[2 * j + 1];
      const int dist = __popc(a.x ^ c.x) + __popc(a.y ^ c.y) + __popc(a.z ^ c.z) + __popc(a.w ^ c.w) + __popc(b.x ^ d.x) + __popc(b.y ^ d.y) +
                       __popc(b.z ^ d.z) + __popc(b.w ^ d.w);      // FORB::distance (FORB.cpp:81-101)
      if (dist < best) { best = dist; best_j = j; }                // first minimum wins (:1244)
    }
    node = V.child_id[best_j];
    if (level == V.nid_level) { nid = node; nid_set = true; }
    j0 = V.child_start[node]; j1 = V.child_start[node + 1];
  } while (j0 != j1);
  if (!nid_set) nid = node;     // leaf above level L - levelsup: *nid is left uninitialised by the reference; declared: the leaf
  const double w = V.weight[node];
  word_id[o] = w > 0 ? V.word_of[node] : -1;                       // "not stopped" (:1157)
  node_id[o] = w > 0 ? nid : -1;
  wgt[o] = w;
}

// ascending bitonic sort of N (power of two) 64-bit keys in shared memory by the whole CTA
__device__ __forceinline__ void bitonic_sort(unsigned long long* key, int N) {
  for (int k = 2; k <= N; k <<= 1)
    for (int j = k >> 1; j > 0; j >>= 1) {
      for (int t = threadIdx.x; t < N; t += blockDim.x) {
        const int p = t ^ j;
        if (p > t) {
          const unsigned long long x = key[t], y = key[p];
          if (((t & k) == 0) == (x > y)) { key[t] = y; key[p] = x; }
        }
      }
      __syncthreads();
    }
}

// run heads of the sorted keys -> head index of every run; returns the number of runs (to every thread)
// pos[u] = sorted position of the head of run u, pos[nruns] = number of valid keys
__device__ __forceinline__ int run_heads(const unsigned long long* key, int N, int* pos, int* s_part) {
  const int tid = threadIdx.x, per = N / blockDim.x;   // N >= blockDim.x, both powers of two
  int c = 0;
  for (int t = tid * per; t < (tid + 1) * per; ++t) {
    const unsigned long long x = key[t];
    if (x != ~0ull && (t == 0 || (unsigned)(key[t - 1] >> 32) != (unsigned)(x >> 32))) ++c;
  }
  s_part[tid] = c;
  __syncthreads();
  if (tid == 0) { int run = 0; for (int i = 0; i < (int)blockDim.x; ++i) { const int t = s_part[i]; s_part[i] = run; run += t; } s_part[blockDim.x] = run; }
  __syncthreads();
  int u = s_part[tid];
  for (int t = tid * per; t < (tid + 1) * per; ++t) {
    const unsigned long long x = key[t];
    if (x != ~0ull && (t == 0 || (unsigned)(key[t - 1] >> 32) != (unsigned)(x >> 32))) pos[u++] = t;
    if (x != ~0ull && (t == N - 1 || key[t + 1] == ~0ull)) pos[s_part[blockDim.x]] = t + 1;
  }
  const int n = s_part[blockDim.x];
  if (n == 0 && tid == 0) pos[0] = 0;
  __syncthreads();
  return n;
}

struct BowOut {
  int* bow_n; int* bow_word; double* bow_value; int* fv_n; int* fv_node; int* fv_start; int* fv_feat;
};
// weighting: 0 TF_IDF, 1 TF (addWeight), 2 IDF, 3 BINARY (addIfNotExist); norm: 0 none (DOT_PRODUCT), 1 L1, 2 L2
__global__ void __launch_bounds__(256) k_bow_reduce(const int* __restrict__ cnt, int cap, int N, const int* __restrict__ word_id,
                                                    const int* __restrict__ node_id, const double* __restrict__ wgt, int weighting, int norm,
                                                    BowOut O) {
  extern __shared__ unsigned long long s_key[];     // [N] keys, then [N + 1] run positions (int)
  __shared__ int s_part[257];
  __shared__ double s_norm;
  int* pos = reinterpret_cast<int*>(s_key + N);
  const int f = blockIdx.x, tid = threadIdx.x;
  const int n = min(cnt[f], cap);
  const long long o = (long long)f * cap;
  // ---- BowVector
  for (int t = tid; t < N; t += 256) {
    const int w = t < n ? word_id[o + t] : -1;
    s_key[t] = w >= 0 ? ((unsigned long long)(unsigned)w << 32 | (unsigned)t) : ~0ull;
  }
  __syncthreads();
  bitonic_sort(s_key, N);
  const int nw = run_heads(s_key, N, pos, s_part);
  double* val = O.bow_value + o;
  for (int u = tid; u < nw; u += 256) {
    const int t0 = pos[u], t1 = pos[u + 1];
    const double w = wgt[o + (unsigned)(s_key[t0] & 0xffffffffu)];
    double v = w;                                                   // insert(id, w)
    if (weighting <= 1) for (int t = t0 + 1; t < t1; ++t) v += w;   // vit->second += v, once per further descriptor
    O.bow_word[o + u] = (int)(s_key[t0] >> 32);
    val[u] = v;
  }
  __syncthreads();
  if (tid == 0) {
    double s = 0.0;
    if (norm == 1) for (int u = 0; u < nw; ++u) s += fabs(val[u]);
    else if (norm == 2) { for (int u = 0; u < nw; ++u) s += val[u] * val[u]; s = sqrt(s); }
    else s = (weighting <= 1 && nw > 0) ? (double)nw : 0.0;         // !must: vit->second /= v.size() (:1164-1170)
    s_norm = s;
    O.bow_n[f] = nw;
  }
  __syncthreads();
  if (s_norm > 0.0) for (int u = tid; u < nw; u += 256) val[u] = val[u] / s_norm;
  __syncthreads();
  // ---- FeatureVector
  for (int t = tid; t < N; t += 256) {
    const int w = t < n ? node_id[o + t] : -1;
    s_key[t] = w >= 0 ? ((unsigned long long)(unsigned)w << 32 | (unsigned)t) : ~0ull;
  }
  __syncthreads();
  bitonic_sort(s_key, N);
  const int nn = run_heads(s_key, N, pos, s_part);
  for (int u = tid; u <= nn; u += 256) {
    O.fv_start[(long long)f * (cap + 1) + u] = pos[u];
    if (u < nn) O.fv_node[o + u] = (int)(s_key[pos[u]] >> 32);
  }
  const int nvalid = pos[nn];
  for (int t = tid; t < nvalid; t += 256) O.fv_feat[o + t] = (int)(s_key[t] & 0xffffffffu);
  if (tid == 0) O.fv_n[f] = nn;
}

}  // namespace drfe

using namespace drfe;

struct drfe_vocab {
  int device = 0, k = 0, L = 0, scoring = 0, weighting = 0, nnodes = 0, nwords = 0;
  std::vector<void*> allocs;
  VocabDev dev{};
  // per-call scratch, grown on demand; transform() is const in the reference and shared by its threads, so calls lock
  std::mutex mu;
  int* d_word = nullptr; int* d_node = nullptr; double* d_wgt = nullptr; int* d_out_i = nullptr; double* d_out_d = nullptr;
  size_t scratch_items = 0;
};

template <class T>
static int voc_alloc(drfe_vocab* v, T** p, size_t count) {
  void* q = nullptr;
  DRFE_CUDA(cudaMalloc(&q, std::max<size_t>(count, 1) * sizeof(T)));
  v->allocs.push_back(q);
  *p = (T*)q;
  return DRFE_OK;
}

void drfe_vocab_destroy(drfe_vocab* v) {
  if (!v) return;
  DeviceScope ds(v->device);
  for (void* p : v->allocs) cudaFree(p);
  delete v;
}

int drfe_vocab_words(const drfe_vocab* v) { return v ? v->nwords : 0; }

int drfe_vocab_create(int k, int L, int scoring, int weighting, int nnodes, const int32_t* parent, const uint8_t* is_leaf,
                      const uint8_t* descriptors, const double* weights, int device, drfe_vocab** out) {
  if (!out || !parent || !is_leaf || !descriptors || !weights || nnodes < 1) { set_error("drfe_vocab_create: null argument"); return DRFE_ERR_ARG; }
  *out = nullptr;
  // the loader's own checks (TemplatedVocabulary.h:1359)
  if (k < 0 || k > 20 || L < 1 || L > 10 || scoring < 0 || scoring > 5 || weighting < 0 || weighting > 3) {
    set_error("drfe_vocab_create: not a vocabulary header (k %d L %d scoring %d weighting %d)", k, L, scoring, weighting); return DRFE_ERR_ARG;
  }
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= 0 || device < 0 || device >= ndev) {
    set_error("drfe_vocab_create: no usable CUDA device %d (no CPU fallback)", device); return DRFE_ERR_CUDA;
  }
  const int total = nnodes + 1;                       // + root
  std::vector<int> nchild(total + 1, 0);
  for (int i = 0; i < nnodes; ++i) {
    if (parent[i] < 0 || parent[i] > i) { set_error("drfe_vocab_create: node %d has parent %d", i + 1, parent[i]); return DRFE_ERR_ARG; }
    ++nchild[parent[i]];
  }
  std::vector<int> start(total + 1, 0), fill(total, 0), child_id(nnodes), word_of(total, 0);
  for (int id = 0; id < total; ++id) start[id + 1] = start[id] + nchild[id];
  std::vector<uint8_t> cdesc((size_t)nnodes * 32);
  std::vector<double> wt(total, 0.0);
  int nwords = 0;
  for (int i = 0; i < nnodes; ++i) {                  // children.push_back(nid) in file order (:1391)
    const int nid = i + 1, slot = start[parent[i]] + fill[parent[i]]++;
    child_id[slot] = nid;
    memcpy(&cdesc[(size_t)slot * 32], descriptors + (size_t)i * 32, 32);
    wt[nid] = weights[i];
    if (is_leaf[i]) word_of[nid] = nwords++;          // :1408-1415
  }
  if (nchild[0] == 0) { set_error("drfe_vocab_create: the root has no children"); return DRFE_ERR_ARG; }
  drfe_vocab* v = new drfe_vocab;
  v->device = device; v->k = k; v->L = L; v->scoring = scoring; v->weighting = weighting; v->nnodes = nnodes; v->nwords = nwords;
  DeviceScope ds(device);
  int* d_start = nullptr; int* d_cid = nullptr; uint4* d_cdesc = nullptr; int* d_word = nullptr; double* d_wt = nullptr;
  auto fail = [&](int rc) { drfe_vocab_destroy(v); return rc; };
  if (voc_alloc(v, &d_start, (size_t)total + 1) || voc_alloc(v, &d_cid, (size_t)nnodes) || voc_alloc(v, &d_cdesc, (size_t)nnodes * 2) ||
      voc_alloc(v, &d_word, (size_t)total) || voc_alloc(v, &d_wt, (size_t)total))
    return fail(DRFE_ERR_CUDA);
  if (cudaMemcpy(d_start, start.data(), ((size_t)total + 1) * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_cid, child_id.data(), (size_t)nnodes * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_cdesc, cdesc.data(), (size_t)nnodes * 32, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_word, word_of.data(), (size_t)total * sizeof(int), cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(d_wt, wt.data(), (size_t)total * sizeof(double), cudaMemcpyHostToDevice) != cudaSuccess) {
    set_error("drfe_vocab_create: upload failed: %s", cudaGetErrorString(cudaGetLastError()));
    return fail(DRFE_ERR_CUDA);
  }
  v->dev.child_start = d_start; v->dev.child_id = d_cid; v->dev.child_desc = d_cdesc; v->dev.word_of = d_word; v->dev.weight = d_wt;
  *out = v;
  return DRFE_OK;
}

int drfe_orb_compute_bow(drfe_orb* h, const drfe_vocab* vc, int levelsup, int32_t* word_id, int32_t* node_id, int* bow_n,
                         int32_t* bow_word, double* bow_value, int* fv_n, int32_t* fv_node, int32_t* fv_start, int32_t* fv_feat) {
  drfe_vocab* v = const_cast<drfe_vocab*>(vc);
  OrbBatchView B;
  if (!v || orb_batch_view(h, &B)) { set_error("drfe_orb_compute_bow: null argument"); return DRFE_ERR_ARG; }
  if (!B.pending) { set_error("drfe_orb_compute_bow: nothing enqueued"); return DRFE_ERR_STATE; }
  if (B.device != v->device) { set_error("drfe_orb_compute_bow: the vocabulary lives on device %d, the extractor on %d", v->device, B.device); return DRFE_ERR_ARG; }
  DeviceScope ds(B.device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  std::lock_guard<std::mutex> lock(v->mu);
  const int nf = B.nframes, cap = B.cap;
  const size_t items = (size_t)nf * cap;
  if (items > v->scratch_items) {
    if (voc_alloc(v, &v->d_word, items) || voc_alloc(v, &v->d_node, items) || voc_alloc(v, &v->d_wgt, items) ||
        voc_alloc(v, &v->d_out_i, items * 3 + (size_t)nf * (cap + 1) + 2 * (size_t)nf) || voc_alloc(v, &v->d_out_d, items))
      return DRFE_ERR_CUDA;
    v->scratch_items = items;
  }
  int N = 256;
  while (N < cap) N <<= 1;
  const size_t smem = (size_t)N * 8 + ((size_t)N + 1) * 4;
  if (smem > 200 * 1024) { set_error("drfe_orb_compute_bow: %d keypoints per frame do not fit", cap); return DRFE_ERR_CAPACITY; }
  DRFE_CUDA(cudaFuncSetAttribute(k_bow_reduce, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaStream_t st = B.stream;
  VocabDev V = v->dev;
  V.nid_level = v->L - levelsup;
  BowOut O;
  O.bow_word = v->d_out_i; O.fv_node = v->d_out_i + items; O.fv_feat = v->d_out_i + 2 * items; O.fv_start = v->d_out_i + 3 * items;
  O.bow_n = O.fv_start + (size_t)nf * (cap + 1); O.fv_n = O.bow_n + nf; O.bow_value = v->d_out_d;
  DRFE_LAUNCH(k_bow_descend, dim3((cap + 127) / 128, nf), 128, 0, st, V, B.desc, B.cnt, cap, v->d_word, v->d_node, v->d_wgt);
  const int norm = (v->scoring == 5) ? 0 : (v->scoring == 1) ? 2 : 1;   // ScoringObject.h:73-89
  DRFE_LAUNCH(k_bow_reduce, nf, 256, smem, st, B.cnt, cap, N, v->d_word, v->d_node, v->d_wgt, v->weighting, norm, O);
  auto d2h = [&](void* dst, const void* src, size_t bytes) { return dst ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, st) : cudaSuccess; };
  DRFE_CUDA(d2h(word_id, v->d_word, items * 4));
  DRFE_CUDA(d2h(node_id, v->d_node, items * 4));
  DRFE_CUDA(d2h(bow_n, O.bow_n, (size_t)nf * 4));
  DRFE_CUDA(d2h(bow_word, O.bow_word, items * 4));
  DRFE_CUDA(d2h(bow_value, O.bow_value, items * 8));
  DRFE_CUDA(d2h(fv_n, O.fv_n, (size_t)nf * 4));
  DRFE_CUDA(d2h(fv_node, O.fv_node, items * 4));
  DRFE_CUDA(d2h(fv_start, O.fv_start, (size_t)nf * (cap + 1) * 4));
  DRFE_CUDA(d2h(fv_feat, O.fv_feat, items * 4));
  DRFE_CUDA(cudaStreamSynchronize(st));
  return DRFE_OK;
}

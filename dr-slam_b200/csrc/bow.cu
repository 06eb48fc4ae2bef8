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

// ------------------------------------------------------------------ ORBmatcher::SearchByBoW(KeyFrame*, Frame&, matches)
// (reference src/ORBmatcher.cc:160-292).  CTA per frame, warp per vocabulary node present in both FeatureVectors.  Inside a
// node the reference walks the keyframe's features in order and skips frame features an earlier one matched (:218-219);
// the warp repeats parallel sweeps against owner[idx] = the earliest position that chose idx in the previous sweep until
// nothing changes, which is the reference's in-order result (see k_search_last_frame in orb.cu).
struct BowSearchDev {
  const int* kf_n; const uint8_t* kf_desc; const float* kf_angle; const uint8_t* kf_valid;
  const int* kf_fv_n; const int* kf_fv_node; const int* kf_fv_start; const int* kf_fv_feat;
  const int* f_fv_n; const int* f_fv_node; const int* f_fv_start; const int* f_fv_feat;
  int* kf_match; int* f_match; int* nmatches;
  int kcap; float nnratio; int check_orientation;
};
__global__ void __launch_bounds__(256) k_search_bow(const uint8_t* __restrict__ fdesc, const drfe_keypoint* __restrict__ fkp,
                                                    const int* __restrict__ fcnt, int cap, BowSearchDev S) {
  extern __shared__ int s_dyn[];
  __shared__ int s_hist[DRFE_HISTO_LENGTH];
  __shared__ int s_keep[3];
  __shared__ int s_cnt[2];
  const int f = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, kcap = S.kcap;
  int* owner = s_dyn;            // [cap]   by frame feature
  int* choice = s_dyn + cap;     // [kcap]  by keyframe feature: bestIdxF if accepted, else -1
  const int nk = min(S.kf_n[f], kcap), nF = min(fcnt[f], cap);
  const uint32_t* FD = reinterpret_cast<const uint32_t*>(fdesc + (long long)f * cap * 32);
  const uint32_t* KD = reinterpret_cast<const uint32_t*>(S.kf_desc + (long long)f * kcap * 32);
  const uint8_t* kvalid = S.kf_valid + (long long)f * kcap;
  const int* knode = S.kf_fv_node + (long long)f * kcap;
  const int* kstart = S.kf_fv_start + (long long)f * (kcap + 1);
  const int* kfeat = S.kf_fv_feat + (long long)f * kcap;
  const int* fnode = S.f_fv_node + (long long)f * cap;
  const int* fstart = S.f_fv_start + (long long)f * (cap + 1);
  const int* ffeat = S.f_fv_feat + (long long)f * cap;
  const int nkn = min(S.kf_fv_n[f], kcap), nfn = min(S.f_fv_n[f], cap);
  for (int i = tid; i < kcap; i += 256) choice[i] = -1;
  for (int i = tid; i < cap; i += 256) owner[i] = 0x7fffffff;
  __syncthreads();
  for (int j = warp; j < nkn; j += 8) {
    // the node of the frame's FeatureVector with the same id (:186-261 is a sorted intersection)
    const int id = knode[j];
    int lo = 0, hi = nfn;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (fnode[mid] < id) lo = mid + 1; else hi = mid; }
    if (lo >= nfn || fnode[lo] != id) continue;
    const int a0 = kstart[j], na = kstart[j + 1] - a0, b0 = fstart[lo], nb = fstart[lo + 1] - b0;
    for (;;) {
      for (int t = lane; t < nb; t += 32) { const int idx = ffeat[b0 + t]; if (idx >= 0 && idx < nF) owner[idx] = 0x7fffffff; }
      __syncwarp();
      for (int t = lane; t < na; t += 32) { const int c = choice[kfeat[a0 + t]]; if (c >= 0) atomicMin(&owner[c], t); }
      __syncwarp();
      int changed = 0;
      for (int t = lane; t < na; t += 32) {
        const int ik = kfeat[a0 + t];
        int c = -1;
        if (ik >= 0 && ik < nk && kvalid[ik]) {                                   // :196-202
          uint32_t d[8];
#pragma unroll
          for (int k = 0; k < 8; ++k) d[k] = KD[ik * 8 + k];
          int best1 = 256, best2 = 256, bidx = -1;
          for (int u = 0; u < nb; ++u) {
            const int idx = ffeat[b0 + u];
            if (idx < 0 || idx >= nF) continue;
            if (owner[idx] < t) continue;                                          // :218-219
            const uint4 p = *reinterpret_cast<const uint4*>(FD + idx * 8), q = *reinterpret_cast<const uint4*>(FD + idx * 8 + 4);
            const int dist = __popc(d[0] ^ p.x) + __popc(d[1] ^ p.y) + __popc(d[2] ^ p.z) + __popc(d[3] ^ p.w) + __popc(d[4] ^ q.x) +
                             __popc(d[5] ^ q.y) + __popc(d[6] ^ q.z) + __popc(d[7] ^ q.w);
            if (dist < best1) { best2 = best1; best1 = dist; bidx = idx; }
            else if (dist < best2) best2 = dist;
          }
          if (best1 <= DRFE_TH_LOW && (float)best1 < __fmul_rn(S.nnratio, (float)best2)) c = bidx;   // :234-238
        }
        changed |= (c != choice[ik >= 0 && ik < kcap ? ik : 0]);
        if (ik >= 0 && ik < kcap) choice[ik] = c;
      }
      if (!__any_sync(0xffffffffu, changed)) break;
      __syncwarp();
    }
  }
  __syncthreads();
  // rotation consistency (:244-254, 271-289)
  if (tid < DRFE_HISTO_LENGTH) s_hist[tid] = 0;
  if (tid < 2) s_cnt[tid] = 0;
  for (int i = tid; i < cap; i += 256) owner[i] = -1;                            // becomes f_match
  __syncthreads();
  const float factor = 1.0f / DRFE_HISTO_LENGTH;
  const float* kang = S.kf_angle + (long long)f * kcap;
  const drfe_keypoint* kp = fkp + (long long)f * cap;
  auto rot_bin = [&](int i) {
    float rot = __fsub_rn(kang[i], kp[choice[i]].angle);
    if (rot < 0.f) rot = __fadd_rn(rot, 360.0f);
    int bin = (int)roundf(__fmul_rn(rot, factor));
    if (bin == DRFE_HISTO_LENGTH) bin = 0;
    return bin;
  };
  int mine = 0;
  for (int i = tid; i < nk; i += 256)
    if (choice[i] >= 0) {
      ++mine;
      owner[choice[i]] = i;                                                        // a frame feature is matched at most once
      if (S.check_orientation) { const int b = rot_bin(i); if (b >= 0 && b < DRFE_HISTO_LENGTH) atomicAdd(&s_hist[b], 1); }
    }
  if (mine) atomicAdd(&s_cnt[0], mine);
  __syncthreads();
  if (tid == 0) {
    int ind1 = -1, ind2 = -1, ind3 = -1;
    if (S.check_orientation) {                                                    // ComputeThreeMaxima (:1666-1707)
      int max1 = 0, max2 = 0, max3 = 0;
      for (int i = 0; i < DRFE_HISTO_LENGTH; ++i) {
        const int s = s_hist[i];
        if (s > max1) { max3 = max2; max2 = max1; max1 = s; ind3 = ind2; ind2 = ind1; ind1 = i; }
        else if (s > max2) { max3 = max2; max2 = s; ind3 = ind2; ind2 = i; }
        else if (s > max3) { max3 = s; ind3 = i; }
      }
      if ((float)max2 < __fmul_rn(0.1f, (float)max1)) { ind2 = -1; ind3 = -1; }
      else if ((float)max3 < __fmul_rn(0.1f, (float)max1)) ind3 = -1;
    }
    s_keep[0] = ind1; s_keep[1] = ind2; s_keep[2] = ind3;
  }
  __syncthreads();
  if (S.check_orientation) {
    int removed = 0;
    for (int i = tid; i < nk; i += 256)
      if (choice[i] >= 0) {
        const int b = rot_bin(i);
        if (b != s_keep[0] && b != s_keep[1] && b != s_keep[2]) { owner[choice[i]] = -1; ++removed; }
      }
    if (removed) atomicAdd(&s_cnt[1], removed);
  }
  __syncthreads();
  for (int i = tid; i < cap; i += 256) S.f_match[(long long)f * cap + i] = owner[i];
  for (int i = tid; i < kcap; i += 256) S.kf_match[(long long)f * kcap + i] = choice[i];
  if (tid == 0) S.nmatches[f] = s_cnt[0] - s_cnt[1];
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
  NvtxRange nvtx_("drfe_orb_compute_bow");
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
  DRFE_CUDA(raise_dyn_smem(k_bow_reduce, B.device, (size_t)(smem)));
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

int drfe_orb_search_by_bow(drfe_orb* h, int kcap, const int* kf_n, const uint8_t* kf_desc, const float* kf_angle, const uint8_t* kf_valid,
                           const int* kf_fv_n, const int32_t* kf_fv_node, const int32_t* kf_fv_start, const int32_t* kf_fv_feat,
                           const int* f_fv_n, const int32_t* f_fv_node, const int32_t* f_fv_start, const int32_t* f_fv_feat, float nnratio,
                           int check_orientation, int32_t* kf_match, int32_t* f_match, int* nmatches) {
  NvtxRange nvtx_("drfe_orb_search_by_bow");
  OrbBatchView B;
  if (orb_batch_view(h, &B) || kcap < 1 || !kf_n || !kf_desc || !kf_angle || !kf_valid || !kf_fv_n || !kf_fv_node || !kf_fv_start ||
      !kf_fv_feat || !f_fv_n || !f_fv_node || !f_fv_start || !f_fv_feat) {
    set_error("drfe_orb_search_by_bow: bad argument"); return DRFE_ERR_ARG;
  }
  if (!B.pending) { set_error("drfe_orb_search_by_bow: nothing enqueued"); return DRFE_ERR_STATE; }
  DeviceScope ds(B.device);
  if (!ds.ok) { set_error("cudaSetDevice failed"); return DRFE_ERR_CUDA; }
  const int nf = B.nframes, cap = B.cap;
  for (int f = 0; f < nf; ++f) {
    if (kf_n[f] < 0 || kf_n[f] > kcap || kf_fv_n[f] < 0 || kf_fv_n[f] > kcap || f_fv_n[f] < 0 || f_fv_n[f] > cap) {
      set_error("drfe_orb_search_by_bow: frame %d: counts out of range", f); return DRFE_ERR_ARG;
    }
    // the lists a FeatureVector can hold: ascending node ids, starts ascending within the feature array
    for (int j = 0; j < kf_fv_n[f]; ++j) {
      const int32_t* st = kf_fv_start + (size_t)f * (kcap + 1);
      if (st[j] < 0 || st[j + 1] < st[j] || st[j + 1] > kcap || (j && kf_fv_node[(size_t)f * kcap + j] <= kf_fv_node[(size_t)f * kcap + j - 1])) {
        set_error("drfe_orb_search_by_bow: frame %d: keyframe FeatureVector entry %d is malformed", f, j); return DRFE_ERR_ARG;
      }
    }
    for (int j = 0; j < f_fv_n[f]; ++j) {
      const int32_t* st = f_fv_start + (size_t)f * (cap + 1);
      if (st[j] < 0 || st[j + 1] < st[j] || st[j + 1] > cap || (j && f_fv_node[(size_t)f * cap + j] <= f_fv_node[(size_t)f * cap + j - 1])) {
        set_error("drfe_orb_search_by_bow: frame %d: frame FeatureVector entry %d is malformed", f, j); return DRFE_ERR_ARG;
      }
    }
  }
  const size_t smem = ((size_t)cap + kcap) * sizeof(int);
  if (smem > 200 * 1024) { set_error("drfe_orb_search_by_bow: kcap %d too large", kcap); return DRFE_ERR_CAPACITY; }
  cudaStream_t st = B.stream;
  // one stream-ordered scratch block per call
  const size_t nk = (size_t)nf * kcap, nc = (size_t)nf * cap;
  size_t off = 0;
  auto take = [&](size_t bytes) { const size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
  const size_t o_desc = take(nk * 32), o_ang = take(nk * 4), o_val = take(nk), o_kn = take((size_t)nf * 4), o_kfn = take((size_t)nf * 4),
               o_knode = take(nk * 4), o_kstart = take((size_t)nf * (kcap + 1) * 4), o_kfeat = take(nk * 4), o_ffn = take((size_t)nf * 4),
               o_fnode = take(nc * 4), o_fstart = take((size_t)nf * (cap + 1) * 4), o_ffeat = take(nc * 4), o_km = take(nk * 4),
               o_fm = take(nc * 4), o_nm = take((size_t)nf * 4);
  char* d = nullptr;
  DRFE_CUDA(cudaMallocAsync((void**)&d, off, st));
  auto up = [&](size_t o, const void* src, size_t bytes) { return cudaMemcpyAsync(d + o, src, bytes, cudaMemcpyHostToDevice, st); };
  cudaError_t e = cudaSuccess;
  auto acc = [&](cudaError_t r) { if (e == cudaSuccess) e = r; };
  acc(up(o_desc, kf_desc, nk * 32)); acc(up(o_ang, kf_angle, nk * 4)); acc(up(o_val, kf_valid, nk)); acc(up(o_kn, kf_n, (size_t)nf * 4));
  acc(up(o_kfn, kf_fv_n, (size_t)nf * 4)); acc(up(o_knode, kf_fv_node, nk * 4)); acc(up(o_kstart, kf_fv_start, (size_t)nf * (kcap + 1) * 4));
  acc(up(o_kfeat, kf_fv_feat, nk * 4)); acc(up(o_ffn, f_fv_n, (size_t)nf * 4)); acc(up(o_fnode, f_fv_node, nc * 4));
  acc(up(o_fstart, f_fv_start, (size_t)nf * (cap + 1) * 4)); acc(up(o_ffeat, f_fv_feat, nc * 4));
  BowSearchDev S;
  S.kf_n = (const int*)(d + o_kn); S.kf_desc = (const uint8_t*)(d + o_desc); S.kf_angle = (const float*)(d + o_ang); S.kf_valid = (const uint8_t*)(d + o_val);
  S.kf_fv_n = (const int*)(d + o_kfn); S.kf_fv_node = (const int*)(d + o_knode); S.kf_fv_start = (const int*)(d + o_kstart); S.kf_fv_feat = (const int*)(d + o_kfeat);
  S.f_fv_n = (const int*)(d + o_ffn); S.f_fv_node = (const int*)(d + o_fnode); S.f_fv_start = (const int*)(d + o_fstart); S.f_fv_feat = (const int*)(d + o_ffeat);
  S.kf_match = (int*)(d + o_km); S.f_match = (int*)(d + o_fm); S.nmatches = (int*)(d + o_nm);
  S.kcap = kcap; S.nnratio = nnratio; S.check_orientation = check_orientation;
  if (e == cudaSuccess) e = raise_dyn_smem(k_search_bow, B.device, (size_t)(smem));
  if (e == cudaSuccess) {
    k_search_bow<<<nf, 256, smem, st>>>(B.desc, B.kp, B.cnt, cap, S);
    g_launches.fetch_add(1, std::memory_order_relaxed);
    e = cudaGetLastError();
  }
  auto down = [&](void* dst, size_t o, size_t bytes) { if (dst && e == cudaSuccess) e = cudaMemcpyAsync(dst, d + o, bytes, cudaMemcpyDeviceToHost, st); };
  down(kf_match, o_km, nk * 4); down(f_match, o_fm, nc * 4); down(nmatches, o_nm, (size_t)nf * 4);
  cudaFreeAsync(d, st);
  if (e == cudaSuccess) e = cudaStreamSynchronize(st);
  if (e != cudaSuccess) { set_error("drfe_orb_search_by_bow: %s", cudaGetErrorString(e)); return DRFE_ERR_CUDA; }
  return DRFE_OK;
}

// Planar_SLAM::ORBVocabulary (reference include/ORBVocabulary.h: DBoW2::TemplatedVocabulary<FORB::TDescriptor, FORB>) on the
// drfe C ABI, for the calls the front end makes: loadFromTextFile (Thirdparty/DBoW2/DBoW2/TemplatedVocabulary.h:1338-1422,
// same file format, same checks) and transform(features, BowVector&, FeatureVector&, levelsup) (:1126-1194) — here on the
// descriptors the extractor left on the device, which is what Frame::ComputeBoW passes (Frame.cc:828-833).
#pragma once
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "ORBextractor.h"

namespace DBoW2 {
typedef unsigned int WordId;
typedef double WordValue;
typedef unsigned int NodeId;
// BowVector.h:56-57 / FeatureVector.h:25-26 (the map interfaces Tracking and the matchers iterate)
class BowVector : public std::map<WordId, WordValue> {};
class FeatureVector : public std::map<NodeId, std::vector<unsigned int> > {};
}  // namespace DBoW2

namespace Planar_SLAM {

class ORBVocabulary {
 public:
  explicit ORBVocabulary(int device = 0) : device_(device) {}
  ~ORBVocabulary() { drfe_vocab_destroy(v_); }
  ORBVocabulary(const ORBVocabulary&) = delete;
  ORBVocabulary& operator=(const ORBVocabulary&) = delete;

  // TemplatedVocabulary::loadFromTextFile: "k L scoring weighting" then one line per node: parent is_leaf 32 bytes weight
  bool loadFromTextFile(const std::string& filename) {
    std::ifstream f(filename.c_str());
    if (!f.is_open()) return false;
    std::string s;
    std::getline(f, s);
    std::stringstream ss(s);
    int k = 0, L = 0, n1 = 0, n2 = 0;
    ss >> k >> L >> n1 >> n2;
    if (k < 0 || k > 20 || L < 1 || L > 10 || n1 < 0 || n1 > 5 || n2 < 0 || n2 > 3) return false;   // :1359-1363
    std::vector<int32_t> parent;
    std::vector<uint8_t> leaf, desc;
    std::vector<double> weight;
    while (std::getline(f, s)) {
      if (s.empty()) continue;
      std::stringstream sn(s);
      int pid = 0, is_leaf = 0;
      sn >> pid >> is_leaf;
      for (int i = 0; i < 32; ++i) { int b = 0; sn >> b; desc.push_back((uint8_t)b); }   // FORB::fromString (FORB.cpp:123-145)
      double w = 0;
      sn >> w;
      if (sn.fail()) return false;
      parent.push_back(pid); leaf.push_back(is_leaf > 0); weight.push_back(w);
    }
    drfe_vocab_destroy(v_);
    v_ = nullptr;
    if (drfe_vocab_create(k, L, n1, n2, (int)parent.size(), parent.data(), leaf.data(), desc.data(), weight.data(), device_, &v_) != DRFE_OK)
      throw std::runtime_error(std::string("ORBVocabulary: ") + drfe_last_error());
    return true;
  }
  unsigned int size() const { return (unsigned)drfe_vocab_words(v_); }
  bool empty() const { return v_ == nullptr || size() == 0; }

  // transform(vCurrentDesc, mBowVec, mFeatVec, levelsup) for the descriptors of the extractor's last call
  void transform(ORBextractor& ex, DBoW2::BowVector& v, DBoW2::FeatureVector& fv, int levelsup) {
    v.clear();
    fv.clear();
    if (empty()) return;                                                                   // :1134-1137
    const int cap = ex.max_keypoints();
    std::vector<int32_t> bw(cap), fnode(cap), fstart(cap + 1), ffeat(cap);
    std::vector<double> bv(cap);
    int bn = 0, fn = 0;
    if (drfe_orb_compute_bow(ex.handle(), v_, levelsup, nullptr, nullptr, &bn, bw.data(), bv.data(), &fn, fnode.data(), fstart.data(), ffeat.data()) != DRFE_OK)
      throw std::runtime_error(std::string("ORBVocabulary::transform: ") + drfe_last_error());
    for (int j = 0; j < bn; ++j) v.insert(v.end(), std::make_pair((DBoW2::WordId)bw[j], bv[j]));          // already in key order
    for (int j = 0; j < fn; ++j)
      fv.insert(fv.end(), std::make_pair((DBoW2::NodeId)fnode[j], std::vector<unsigned int>(ffeat.begin() + fstart[j], ffeat.begin() + fstart[j + 1])));
  }
  drfe_vocab* handle() const { return v_; }

 private:
  int device_ = 0;
  drfe_vocab* v_ = nullptr;
};

}  // namespace Planar_SLAM

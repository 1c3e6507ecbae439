// Batch containers of the hot path + the synthetic parallel-corpus generator.
//
// SubBatch / CorpusBatch reproduce the layout of the reference
// (src/data/corpus.h:49-205): per side, time-major `indices[t*B + b]` and
// `mask[t*B + b]` in {0,1}; split(n) cuts contiguous sentence ranges of
// ceil(B/n) for data parallelism (:73-102,145-169); words() counts the SOURCE
// side only (:173).  Text I/O (Corpus, Vocab, BatchGenerator) is out of scope;
// batches come from SyntheticCorpus below.
#pragma once

#include <algorithm>
#include <cmath>
#include <random>
#include <vector>

#include "common/definitions.h"

namespace marian {

typedef size_t Word;
const Word EOS_ID = 0;
const Word UNK_ID = 1;

namespace data {

class SubBatch {
public:
  SubBatch(int size, int width) : indices_((size_t)size * width, 0), mask_((size_t)size * width, 0), size_(size), width_(width), words_(0) {}

  std::vector<Word>& indices() { return indices_; }
  std::vector<float>& mask() { return mask_; }
  const std::vector<Word>& indices() const { return indices_; }
  const std::vector<float>& mask() const { return mask_; }

  size_t batchSize() const { return size_; }
  size_t batchWidth() const { return width_; }
  size_t batchWords() const { return words_; }
  void setWords(size_t words) { words_ = words; }

  std::vector<Ptr<SubBatch>> split(size_t n) {
    std::vector<Ptr<SubBatch>> splits;
    size_t subSize = (size_t)std::ceil(size_ / (float)n);
    size_t totSize = size_;
    size_t pos = 0;
    for(size_t k = 0; k < n; ++k) {
      size_t sz = std::min(subSize, totSize);
      auto sb = New<SubBatch>((int)sz, (int)width_);
      size_t words = 0;
      for(size_t j = 0; j < width_; ++j)
        for(size_t i = 0; i < sz; ++i) {
          sb->indices()[j * sz + i] = indices_[j * size_ + pos + i];
          sb->mask()[j * sz + i] = mask_[j * size_ + pos + i];
          if(mask_[j * size_ + pos + i] != 0)
            words++;
        }
      sb->setWords(words);
      splits.push_back(sb);
      totSize -= sz;
      pos += sz;
    }
    return splits;
  }

private:
  std::vector<Word> indices_;
  std::vector<float> mask_;
  size_t size_;
  size_t width_;
  size_t words_;
};

class CorpusBatch {
public:
  explicit CorpusBatch(const std::vector<Ptr<SubBatch>>& batches) : batches_(batches) {}

  Ptr<SubBatch> operator[](size_t i) const { return batches_[i]; }
  Ptr<SubBatch> front() { return batches_.front(); }
  Ptr<SubBatch> back() { return batches_.back(); }

  size_t size() const { return batches_[0]->batchSize(); }
  size_t words() const { return batches_[0]->batchWords(); }  // source side only, as the reference logs
  size_t wordsTotal() const {
    size_t w = 0;
    for(auto& b : batches_)
      w += b->batchWords();
    return w;
  }
  size_t sets() const { return batches_.size(); }

  // line numbers of the sentences in their corpus (translations are written back in corpus order; reference
  // src/data/batch.h getSentenceIds).  Empty for synthetic batches: sentence i is line i.
  const std::vector<size_t>& getSentenceIds() const { return sentenceIds_; }
  void setSentenceIds(const std::vector<size_t>& ids) { sentenceIds_ = ids; }

  std::vector<Ptr<CorpusBatch>> split(size_t n) {
    std::vector<std::vector<Ptr<SubBatch>>> subs(n);
    for(auto subBatch : batches_) {
      size_t i = 0;
      for(auto s : subBatch->split(n))
        subs[i++].push_back(s);
    }
    std::vector<Ptr<CorpusBatch>> splits;
    for(auto& s : subs)
      splits.push_back(New<CorpusBatch>(s));
    return splits;
  }

  // Shape key: two batches with equal keys build identical tapes.
  std::vector<int> shapeKey() const {
    std::vector<int> k;
    for(auto& b : batches_) {
      k.push_back((int)b->batchSize());
      k.push_back((int)b->batchWidth());
    }
    return k;
  }

  static Ptr<CorpusBatch> fakeBatch(const std::vector<size_t>& lengths, size_t batchSize) {
    std::vector<Ptr<SubBatch>> batches;
    for(auto len : lengths) {
      auto sb = New<SubBatch>((int)batchSize, (int)len);
      std::fill(sb->mask().begin(), sb->mask().end(), 1.f);
      sb->setWords(batchSize * len);
      batches.push_back(sb);
    }
    return New<CorpusBatch>(batches);
  }

private:
  std::vector<Ptr<SubBatch>> batches_;
  std::vector<size_t> sentenceIds_;
};

// Synthetic bitext (SURVEY.md 8d / BASELINE.md section 3): a sentence is len-1
// ids uniform in [2, V) from std::mt19937(seed) followed by EOS (id 0), which
// the reference appends and counts (src/data/vocab.cpp:31-32).  "dense":
// every sentence has exactly T tokens; "padded": lengths uniform in
// [T/2, T], batch padded to its maximum and sorted by target length
// (descending) as maxi-batch-sort trg would (src/data/batch_generator.h:45-62).
class SyntheticCorpus {
public:
  SyntheticCorpus(int vocabSrc, int vocabTrg, uint32_t seed = 1111) : vs_(vocabSrc), vt_(vocabTrg), rng_(seed) {}

  Ptr<CorpusBatch> next(int batchSize, int maxLenSrc, int maxLenTrg, bool padded) {
    std::vector<std::vector<Word>> src(batchSize), trg(batchSize);
    for(int b = 0; b < batchSize; ++b) {
      src[b] = sentence(length(maxLenSrc, padded), vs_);
      trg[b] = sentence(length(maxLenTrg, padded), vt_);
    }
    if(padded) {
      std::vector<int> order(batchSize);
      for(int i = 0; i < batchSize; ++i)
        order[i] = i;
      std::stable_sort(order.begin(), order.end(), [&](int a, int b) { return trg[a].size() > trg[b].size(); });
      std::vector<std::vector<Word>> s2, t2;
      for(int i : order) {
        s2.push_back(src[i]);
        t2.push_back(trg[i]);
      }
      src.swap(s2);
      trg.swap(t2);
    }
    return New<CorpusBatch>(std::vector<Ptr<SubBatch>>{pack(src), pack(trg)});
  }

private:
  int length(int maxLen, bool padded) {
    if(!padded)
      return maxLen;
    int lo = std::max(1, maxLen / 2);
    return lo + (int)(rng_() % (uint32_t)(maxLen - lo + 1));
  }
  std::vector<Word> sentence(int len, int vocab) {
    std::vector<Word> s(len);
    for(int i = 0; i < len - 1; ++i)
      s[i] = 2 + (Word)(rng_() % (uint32_t)(vocab - 2));
    s[len - 1] = EOS_ID;
    return s;
  }
  static Ptr<SubBatch> pack(const std::vector<std::vector<Word>>& sents) {
    size_t width = 0;
    for(auto& s : sents)
      width = std::max(width, s.size());
    size_t B = sents.size();
    auto sb = New<SubBatch>((int)B, (int)width);
    size_t words = 0;
    for(size_t b = 0; b < B; ++b)
      for(size_t t = 0; t < sents[b].size(); ++t) {
        sb->indices()[t * B + b] = sents[b][t];
        sb->mask()[t * B + b] = 1.f;
        words++;
      }
    sb->setWords(words);
    return sb;
  }

  int vs_, vt_;
  std::mt19937 rng_;
};

}  // namespace data
}  // namespace marian

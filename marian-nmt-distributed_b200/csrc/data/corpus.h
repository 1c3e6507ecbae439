// Text corpus -> mini-batches: the input step in front of the hot path.
//
// Semantics of the reference's data layer, restated for this engine (no Boost, no yaml-cpp):
//   Vocab           src/data/vocab.cpp:16-204   "</s>" = 0, "<unk>" = 1; a vocabulary file is a YAML map
//                   word -> id; create() counts the training file and numbers the words by falling frequency
//                   from 2; operator()(line) splits on blanks, maps unknown words to <unk> and appends </s>.
//   Corpus          src/data/corpus.cpp:30-230  one text file per side; next() yields the next sentence tuple
//                   whose sentences are all non-empty and at most max-length tokens (longer tuples are skipped,
//                   or cropped with max-length-crop); shuffle() permutes the sentence order (seeded mt19937).
//   BatchGenerator  src/data/batch_generator.h:39-160  reads mini-batch x maxi-batch tuples into a priority queue
//                   ordered by source / target length (maxi-batch-sort), pops them into mini-batches of
//                   `mini-batch` sentences (or more than `mini-batch-words` source tokens), shuffles the batches
//                   of the maxi-batch, hands them out one by one.  toBatch() = the reference's time-major
//                   SubBatch layout (src/data/corpus.h:335-376).
// B200-first addition: PrefetchingBatchGenerator runs all of that on a host thread, two mini-batches ahead of
// the device (the reference reads and batches on the training thread between updates); the per-step upload of
// indices / masks already goes through the graph's pinned staging (graph/expression_graph.h batchUploads).
// Not carried over: mini-batch-fit (needs the allocator's fits() probe per length class), guided alignment,
// temporary shuffle files (the shuffle is in memory).
#pragma once

#include <algorithm>
#include <condition_variable>
#include <deque>
#include <fstream>
#include <functional>
#include <map>
#include <mutex>
#include <queue>
#include <random>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "common/options.h"
#include "data/batch.h"

namespace marian {
namespace data {

const std::string EOS_STR = "</s>";
const std::string UNK_STR = "<unk>";

typedef std::vector<Word> Words;

class Vocab {
public:
  size_t operator[](const std::string& word) const {
    auto it = str2id_.find(word);
    return it != str2id_.end() ? it->second : UNK_ID;
  }
  const std::string& operator[](size_t id) const {
    ABORT_IF(id >= id2str_.size(), "Unknown word id:", id);
    return id2str_[id];
  }
  // reference vocab.cpp:25-43
  Words operator()(const std::string& line, bool addEOS = true) const {
    Words words;
    std::istringstream ss(line);
    std::string tok;
    while(ss >> tok)
      words.push_back((*this)[tok]);
    if(addEOS)
      words.push_back(EOS_ID);
    return words;
  }
  std::vector<std::string> operator()(const Words& sentence, bool ignoreEOS = true) const {
    std::vector<std::string> decoded;
    for(auto w : sentence)
      if(w != EOS_ID || !ignoreEOS)
        decoded.push_back((*this)[w]);
    return decoded;
  }
  size_t size() const { return id2str_.size(); }

  // reference vocab.cpp:64-80
  int loadOrCreate(const std::string& vocabPath, const std::string& trainPath, int max) {
    std::string path = vocabPath.empty() ? trainPath + ".yml" : vocabPath;
    if(!std::ifstream(path).good())
      create(path, trainPath);
    return load(path, max);
  }

  // YAML map "word: id", one entry per line; keys may be quoted (yaml-cpp quotes what needs it)
  int load(const std::string& vocabPath, int max = 0) {
    std::ifstream in(vocabPath);
    ABORT_IF(!in.good(), "Vocabulary does not exist:", vocabPath);
    std::string line;
    while(std::getline(in, line)) {
      if(line.empty() || line[0] == '#')
        continue;
      auto colon = line.rfind(':');
      if(colon == std::string::npos)
        continue;
      std::string key = unquote(trim(line.substr(0, colon)));
      std::string val = trim(line.substr(colon + 1));
      if(key.empty() || val.empty())
        continue;
      Word id = (Word)std::stoul(val);
      if(!max || id < (Word)max) {
        str2id_[key] = id;
        if(id >= id2str_.size())
          id2str_.resize(id + 1);
        id2str_[id] = key;
      }
    }
    ABORT_IF(id2str_.empty(), "Empty vocabulary:", vocabPath);
    if(id2str_.size() < 2)
      id2str_.resize(2);
    id2str_[EOS_ID] = EOS_STR;
    id2str_[UNK_ID] = UNK_STR;
    str2id_[EOS_STR] = EOS_ID;
    str2id_[UNK_STR] = UNK_ID;
    return std::max((int)id2str_.size(), max);
  }

  // reference vocab.cpp:151-204: words numbered from 2 by falling frequency (ties: first occurrence first)
  void create(const std::string& vocabPath, const std::string& trainPath, size_t maxSize = 0) {
    std::ifstream in(trainPath);
    ABORT_IF(!in.good(), "Cannot read the training file to create a vocabulary:", trainPath);
    ABORT_IF(std::ifstream(vocabPath).good(), "Vocab file exists. Not overwriting:", vocabPath);
    std::unordered_map<std::string, std::pair<size_t, size_t>> counter;  // word -> (count, first occurrence)
    std::string line, tok;
    size_t order = 0;
    while(std::getline(in, line)) {
      std::istringstream ss(line);
      while(ss >> tok) {
        if(tok == EOS_STR || tok == UNK_STR)
          continue;
        auto it = counter.find(tok);
        if(it == counter.end())
          counter[tok] = {1, order++};
        else
          it->second.first++;
      }
    }
    std::vector<std::string> words;
    for(auto& p : counter)
      words.push_back(p.first);
    std::sort(words.begin(), words.end(), [&](const std::string& a, const std::string& b) {
      auto &ca = counter[a], &cb = counter[b];
      return ca.first != cb.first ? ca.first > cb.first : ca.second < cb.second;
    });
    size_t n = words.size();
    if(maxSize > 2)
      n = std::min(maxSize - 2, n);
    std::ofstream out(vocabPath);
    ABORT_IF(!out.good(), "Cannot write the vocabulary:", vocabPath);
    out << quote(EOS_STR) << ": " << EOS_ID << "\n" << quote(UNK_STR) << ": " << UNK_ID << "\n";
    for(size_t i = 0; i < n; ++i)
      out << quote(words[i]) << ": " << (i + 2) << "\n";
  }

private:
  static std::string trim(const std::string& s) {
    size_t b = s.find_first_not_of(" \t\r"), e = s.find_last_not_of(" \t\r");
    return b == std::string::npos ? "" : s.substr(b, e - b + 1);
  }
  static std::string unquote(const std::string& s) {
    if(s.size() >= 2 && ((s.front() == '"' && s.back() == '"') || (s.front() == '\'' && s.back() == '\''))) {
      std::string out;
      bool dq = s.front() == '"';
      for(size_t i = 1; i + 1 < s.size(); ++i) {
        if(dq && s[i] == '\\' && i + 2 < s.size()) {
          out += s[++i];
        } else if(!dq && s[i] == '\'' && s[i + 1] == '\'' && i + 2 < s.size()) {
          out += '\'';
          ++i;
        } else
          out += s[i];
      }
      return out;
    }
    return s;
  }
  static std::string quote(const std::string& s) {
    std::string out = "\"";
    for(char c : s) {
      if(c == '"' || c == '\\')
        out += '\\';
      out += c;
    }
    return out + "\"";
  }

  std::unordered_map<std::string, size_t> str2id_;
  std::vector<std::string> id2str_;
};

class SentenceTuple {
public:
  explicit SentenceTuple(size_t id = 0) : id_(id) {}
  size_t getId() const { return id_; }
  void push_back(const Words& words) { tuple_.push_back(words); }
  size_t size() const { return tuple_.size(); }
  bool empty() const { return tuple_.empty(); }
  Words& operator[](size_t i) { return tuple_[i]; }
  const Words& operator[](size_t i) const { return tuple_[i]; }
  const Words& back() const { return tuple_.back(); }
  std::vector<Words>::const_iterator begin() const { return tuple_.begin(); }
  std::vector<Words>::const_iterator end() const { return tuple_.end(); }

private:
  size_t id_;
  std::vector<Words> tuple_;
};

class Corpus {
public:
  Corpus(const std::vector<std::string>& paths, const std::vector<Ptr<Vocab>>& vocabs, Ptr<Options> options)
      : paths_(paths),
        vocabs_(vocabs),
        maxLength_(options->get<size_t>("max-length", 50)),
        maxLengthCrop_(options->get<bool>("max-length-crop", false)),
        rightLeft_(options->get<bool>("right-left", false)),
        g_((uint32_t)options->get<size_t>("seed", 1234)) {
    ABORT_IF(paths_.size() != vocabs_.size(), "Number of corpus files and vocab files does not agree");
    for(auto& p : paths_) {
      std::ifstream in(p);
      ABORT_IF(!in.good(), "Cannot read corpus file:", p);
      std::vector<std::string> lines;
      std::string line;
      while(std::getline(in, line))
        lines.push_back(line);
      ABORT_IF(lines.empty(), "File is empty:", p);
      lines_.push_back(std::move(lines));
    }
    size_t n = lines_[0].size();
    for(auto& l : lines_)
      n = std::min(n, l.size());  // the reference stops at the shortest file (corpus.cpp:175-176)
    order_.resize(n);
    reset();
  }

  // reference corpus.cpp:146-187
  SentenceTuple next() {
    while(pos_ < order_.size()) {
      size_t curId = order_[pos_++];
      SentenceTuple tup(curId);
      for(size_t i = 0; i < lines_.size(); ++i) {
        Words words = (*vocabs_[i])(lines_[i][curId]);
        if(words.empty())
          words.push_back(0);
        if(maxLengthCrop_ && words.size() > maxLength_) {
          words.resize(maxLength_);
          words.back() = 0;
        }
        if(rightLeft_)
          std::reverse(words.begin(), words.end() - 1);
        tup.push_back(words);
      }
      if(std::all_of(tup.begin(), tup.end(), [=](const Words& w) { return w.size() > 0 && w.size() <= maxLength_; }))
        return tup;
    }
    return SentenceTuple(0);  // empty: end of the epoch
  }
  void shuffle() {
    reset();
    std::shuffle(order_.begin(), order_.end(), g_);
  }
  void reset() {
    for(size_t i = 0; i < order_.size(); ++i)
      order_[i] = i;
    pos_ = 0;
  }
  size_t sentences() const { return order_.size(); }
  std::vector<Ptr<Vocab>>& getVocabs() { return vocabs_; }

  // reference corpus.h:335-376
  static Ptr<CorpusBatch> toBatch(const std::vector<SentenceTuple>& batchVector) {
    int batchSize = (int)batchVector.size();
    std::vector<int> maxDims;
    for(auto& ex : batchVector) {
      if(maxDims.size() < ex.size())
        maxDims.resize(ex.size(), 0);
      for(size_t i = 0; i < ex.size(); ++i)
        if(ex[i].size() > (size_t)maxDims[i])
          maxDims[i] = (int)ex[i].size();
    }
    std::vector<Ptr<SubBatch>> subBatches;
    for(auto m : maxDims)
      subBatches.emplace_back(New<SubBatch>(batchSize, m));
    std::vector<size_t> words(maxDims.size(), 0);
    for(int i = 0; i < batchSize; ++i)
      for(size_t j = 0; j < maxDims.size(); ++j)
        for(size_t k = 0; k < batchVector[i][j].size(); ++k) {
          subBatches[j]->indices()[k * batchSize + i] = batchVector[i][j][k];
          subBatches[j]->mask()[k * batchSize + i] = 1.f;
          words[j]++;
        }
    for(size_t j = 0; j < maxDims.size(); ++j)
      subBatches[j]->setWords(words[j]);
    auto batch = New<CorpusBatch>(subBatches);
    std::vector<size_t> ids;
    for(auto& ex : batchVector)
      ids.push_back(ex.getId());
    batch->setSentenceIds(ids);
    return batch;
  }

private:
  std::vector<std::string> paths_;
  std::vector<Ptr<Vocab>> vocabs_;
  size_t maxLength_;
  bool maxLengthCrop_, rightLeft_;
  std::mt19937 g_;
  std::vector<std::vector<std::string>> lines_;
  std::vector<size_t> order_;
  size_t pos_{0};
};

// reference batch_generator.h:17-160
class BatchGenerator {
public:
  BatchGenerator(Ptr<Corpus> data, Ptr<Options> options) : data_(data), options_(options), g_((uint32_t)options->get<size_t>("seed", 1234)) {}

  operator bool() const { return !bufferedBatches_.empty(); }

  Ptr<CorpusBatch> next() {
    ABORT_IF(bufferedBatches_.empty(), "No batches to fetch, run prepare()");
    auto current = bufferedBatches_.front();
    bufferedBatches_.pop_front();
    if(bufferedBatches_.empty())
      fillBatches(shuffle_);
    return current;
  }
  void prepare(bool shuffle = true) {
    shuffle_ = shuffle;
    if(shuffle)
      data_->shuffle();
    else
      data_->reset();
    done_ = false;
    lookahead_ = data_->next();
    fillBatches(shuffle);
  }

private:
  typedef SentenceTuple sample;
  void fillBatches(bool shuffle) {
    auto cmpSrc = [](const sample& a, const sample& b) { return a[0].size() < b[0].size(); };
    auto cmpTrg = [](const sample& a, const sample& b) { return a.back().size() < b.back().size(); };
    // "none": the reference compares addresses of queue elements (arbitrary); here arrival order is kept
    auto cmpNone = [](const sample& a, const sample& b) { return a.getId() > b.getId(); };
    typedef std::function<bool(const sample&, const sample&)> cmp_type;
    std::string sort = options_->get<std::string>("maxi-batch-sort", "trg");
    std::priority_queue<sample, std::vector<sample>, cmp_type> maxiBatch(sort == "src" ? cmp_type(cmpSrc) : (sort == "none" ? cmp_type(cmpNone) : cmp_type(cmpTrg)));

    size_t maxBatchSize = (size_t)options_->get<int>("mini-batch", 64);
    size_t maxSize = maxBatchSize * (size_t)options_->get<int>("maxi-batch", 100);
    while(!lookahead_.empty() && maxiBatch.size() < maxSize) {
      maxiBatch.push(lookahead_);
      lookahead_ = data_->next();
    }
    std::vector<sample> batchVector;
    int currentWords = 0;
    int mbWords = options_->get<int>("mini-batch-words", 0);
    while(!maxiBatch.empty()) {
      batchVector.push_back(maxiBatch.top());
      currentWords += (int)batchVector.back()[0].size();
      maxiBatch.pop();
      bool makeBatch = batchVector.size() == maxBatchSize;
      if(mbWords > 0)
        makeBatch = currentWords > mbWords;
      if(makeBatch) {
        bufferedBatches_.push_back(Corpus::toBatch(batchVector));
        batchVector.clear();
        currentWords = 0;
      }
    }
    if(!batchVector.empty())
      bufferedBatches_.push_back(Corpus::toBatch(batchVector));
    if(shuffle)
      std::shuffle(bufferedBatches_.begin(), bufferedBatches_.end(), g_);
  }

  Ptr<Corpus> data_;
  Ptr<Options> options_;
  std::deque<Ptr<CorpusBatch>> bufferedBatches_;
  SentenceTuple lookahead_;
  bool shuffle_{true}, done_{false};
  std::mt19937 g_;
};

// Host thread that keeps `depth` mini-batches ready: tokenisation, length sorting and batch assembly overlap the
// device's work on the previous updates.  next() returns nullptr at the end of an epoch (then restart()).
class PrefetchingBatchGenerator {
public:
  PrefetchingBatchGenerator(Ptr<Corpus> data, Ptr<Options> options, size_t depth = 2) : gen_(data, options), shuffle_(options->get<bool>("shuffle", true)), depth_(depth) {
    start();
  }
  ~PrefetchingBatchGenerator() { stop(); }

  Ptr<CorpusBatch> next() {
    std::unique_lock<std::mutex> lock(m_);
    cv_.wait(lock, [&] { return !ready_.empty() || finished_; });
    if(ready_.empty())
      return nullptr;
    auto b = ready_.front();
    ready_.pop_front();
    cv_.notify_all();
    return b;
  }
  void restart() {
    stop();
    start();
  }

private:
  void start() {
    finished_ = false;
    quit_ = false;
    worker_ = std::thread([this] {
      gen_.prepare(shuffle_);
      while(true) {
        Ptr<CorpusBatch> b = gen_ ? gen_.next() : nullptr;
        std::unique_lock<std::mutex> lock(m_);
        if(!b) {
          finished_ = true;
          cv_.notify_all();
          return;
        }
        cv_.wait(lock, [&] { return ready_.size() < depth_ || quit_; });
        if(quit_)
          return;
        ready_.push_back(b);
        cv_.notify_all();
      }
    });
  }
  void stop() {
    {
      std::unique_lock<std::mutex> lock(m_);
      quit_ = true;
      cv_.notify_all();
    }
    if(worker_.joinable())
      worker_.join();
    ready_.clear();
  }

  BatchGenerator gen_;
  bool shuffle_;
  size_t depth_;
  std::thread worker_;
  std::mutex m_;
  std::condition_variable cv_;
  std::deque<Ptr<CorpusBatch>> ready_;
  bool finished_{false}, quit_{false};
};

}  // namespace data
}  // namespace marian

// ARPA text LM -> LM FSA file (host code, no CUDA): the ingestion step in front of asrd_lm_create.
// Follows the published behaviour of the reference's converter with one thread —
// src/newlm/arpa2fsa.cc:311-739 (AnasyArpa, ReadSymbols, AnalyLine, AddLineToFsa, NgramToFsa,
// ConvertArpa2Fsa) — and writes the file ArpaLm::Read takes (src/newlm/arpa2fsa.h:399-480,
// Fsa::Write arpa2fsa.cc:9-68), byte for byte (tests/test_arpa_ingestion.py holds it against the
// unmodified reference converter built into oracle/_ref).
//
// The FSA (arpa2fsa.h:217-480): state 0 is the start / unigram state, direct-indexed by word id
// (arc k leads to state k + 1); every n-gram line adds one state reached from its history state
// over an arc labelled with the last word; a state's back-off pointer is the state of the longest
// proper suffix of its n-gram that exists.  Weights are natural logs (ARPA log10 x ln 10).
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <unordered_map>
#include <vector>

#include "asrd.h"

namespace {

struct LmArc {
  int32_t word;
  float weight;
  int32_t to;
};

struct LmState {
  std::vector<LmArc> arcs;  // sorted by word (stable insertion, arpa2fsa.h:158-191)
  float backoff_prob = 0.f;
  int32_t backoff_id = 0;
};

const LmArc *FindArc(const LmState &s, int32_t word) {  // FsaState::SearchArc, arpa2fsa.h:194-210
  int lo = 0, hi = (int)s.arcs.size() - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) / 2;
    if (s.arcs[mid].word > word) hi = mid - 1;
    else if (s.arcs[mid].word < word) lo = mid + 1;
    else return &s.arcs[mid];
  }
  return nullptr;
}

void InsertArc(LmState &s, int32_t word, float weight, int32_t to) {
  size_t i = s.arcs.size();
  s.arcs.push_back(LmArc{word, weight, to});
  while (i > 0 && s.arcs[i - 1].word > word) {
    s.arcs[i] = s.arcs[i - 1];
    --i;
  }
  s.arcs[i] = LmArc{word, weight, to};
}

bool SkippedLine(const char *line) {  // arpa2fsa.cc:329-333
  return line[0] == ' ' || line[0] == '\r' || line[0] == '\n' || line[0] == '\t' || line[0] == '#';
}

std::string Upper(const char *s) {
  std::string u(s);
  for (char &c : u) c = (char)toupper((unsigned char)c);
  return u;
}

// log10 -> ln exactly like `float *= M_LN10` (the product is formed in double, rounded once)
float ToLn(float v) { return (float)((double)v * M_LN10); }

struct Line {
  std::vector<int32_t> words;
  float logprob = 0.f, backoff = 0.f;
};

}  // namespace

extern "C" int asrd_lm_convert_arpa(const char *arpa_path, const char *wordlist_path, const char *out_path) {
  if (!arpa_path || !wordlist_path || !out_path) return ASRD_ERR_BAD_ARG;
  // ---- word list: "word id" per line (ArpaLm::ReadSymbols, arpa2fsa.cc:412-447)
  std::unordered_map<std::string, int32_t> syms;
  {
    FILE *fp = fopen(wordlist_path, "r");
    if (!fp) return ASRD_ERR_IO;
    char line[256], word[128];
    int id = 0;
    while (fgets(line, sizeof(line), fp)) {
      if (sscanf(line, "%127s %d", word, &id) != 2) {
        fclose(fp);
        return ASRD_ERR_IO;
      }
      syms[word] = id;
    }
    fclose(fp);
  }
  const int32_t bos = syms["<s>"], eos = syms["</s>"], unk = syms["<unk>"];  // arpa2fsa.cc:699-701
  auto word_id = [&](const char *w) {
    auto it = syms.find(w);
    return it == syms.end() ? unk : it->second;
  };
  // ---- pass 1: the \data\ header and where every n-gram section starts (AnasyArpa, arpa2fsa.cc:311-410)
  std::vector<int32_t> declared;                    // "ngram n=count" as declared
  std::vector<std::vector<std::string>> sections;   // the lines of every \n-grams: section
  {
    FILE *fp = fopen(arpa_path, "r");
    if (!fp) return ASRD_ERR_IO;
    char line[1024];
    enum { kNone, kData, kGram } where = kNone;
    while (fgets(line, sizeof(line), fp)) {
      if (SkippedLine(line)) continue;
      if (line[0] == '\\') {
        const std::string up = Upper(line);
        if (up.find("\\DATA\\") != std::string::npos) where = kData;
        else if (up.find("-GRAMS:") != std::string::npos) {
          where = kGram;
          sections.emplace_back();
        }
        continue;  // (\end\ changes nothing: lines after it would still count for the last section)
      }
      if (where == kData) {
        const std::string up = Upper(line);
        if (up.find("NGRAM") != std::string::npos) {
          int n = 0, count = 0;
          if (sscanf(up.c_str(), "%*s %d=%d", &n, &count) != 2) {
            fclose(fp);
            return ASRD_ERR_IO;
          }
          declared.push_back(count);
        }
      } else if (where == kGram) {
        sections.back().push_back(line);
      }
    }
    fclose(fp);
  }
  if (declared.empty() || sections.size() > declared.size()) return ASRD_ERR_IO;
  const size_t order = declared.size();
  // ---- pass 2: section by section, line by line (NgramToFsa + AddLineToFsa, arpa2fsa.cc:510-683)
  std::vector<LmState> st(1);  // state 0 = start
  auto add_state = [&]() {
    st.emplace_back();
    return (int32_t)st.size() - 1;
  };
  // The reference keeps every state's arcs contiguous in one pool and asserts that a state only
  // grows while it is the last one to have grown (arpa2fsa.h:158-171): n-grams must come grouped by
  // history.  Same requirement here, reported instead of aborting.
  int32_t last_grown = -1;
  auto grow = [&](int32_t s, int32_t word, float w, int32_t to) -> bool {
    if (!st[s].arcs.empty() && last_grown != s) return false;
    InsertArc(st[s], word, w, to);
    last_grown = s;
    return true;
  };
  for (size_t gi = 0; gi < sections.size(); ++gi) {
    const int gram = (int)gi + 1;
    Line prev, cur;
    int32_t hist_state = 0;
    for (const std::string &text : sections[gi]) {
      // ---- AnalyLine (arpa2fsa.cc:450-508)
      cur = Line();
      std::vector<char> buf(text.begin(), text.end());
      buf.push_back('\0');
      bool ok = sscanf(buf.data(), "%f", &cur.logprob) == 1;
      if (ok) {
        cur.logprob = ToLn(cur.logprob);
        char *save = nullptr;
        strtok_r(buf.data(), " \r\n\t", &save);
        for (int i = 0; i < gram && ok; ++i) {
          const char *w = strtok_r(nullptr, " \n\r\t", &save);
          if (!w) {
            ok = false;
            break;
          }
          const int32_t id = word_id(w);
          if (id == unk || (id == bos && i != 0) || (id == eos && i != gram - 1)) {
            ok = false;
            break;
          }
          cur.words.push_back(id);
        }
        if (ok) {
          if ((size_t)gram < order) {
            const char *b = strtok_r(nullptr, " \n\r\t", &save);
            if (b && sscanf(b, "%f", &cur.backoff) != 1) ok = false;
          }
          cur.backoff = ToLn(cur.backoff);
          if (ok && (double)cur.logprob / M_LN10 < -98.0 && !(gram == 1 && cur.words[0] == bos)) ok = false;
        }
      }
      if (!ok) continue;  // the line is dropped; the previous line stays the previous line
      // same history as the previous accepted line?  (ArpaLine::operator==, arpa2fsa.h:531-541)
      bool same = prev.words.size() == cur.words.size();
      for (size_t i = 0; same && i + 1 < cur.words.size(); ++i) same = prev.words[i] == cur.words[i];
      if (!same) hist_state = 0;
      if (gram == 1) {
        const int32_t w = cur.words[0];
        if (w < 0) continue;
        while ((int32_t)st[0].arcs.size() - 1 < w) {  // fill the direct index up to w
          const int32_t id = add_state();
          if (!grow(0, (int32_t)st[0].arcs.size(), 0.0f, id)) return ASRD_ERR_IO;
        }
        LmArc &a = st[0].arcs[w];
        a.weight = cur.logprob;
        st[a.to].backoff_id = 0;
        st[a.to].backoff_prob = cur.backoff;
      } else {
        const size_t n = cur.words.size();
        bool found = true;
        if (hist_state == 0) {  // walk the history from the start state
          int32_t s = 0;
          for (size_t i = 0; i + 1 < n; ++i) {
            const LmArc *a = nullptr;
            if (i == 0) {
              if (cur.words[0] < (int32_t)st[0].arcs.size()) a = &st[0].arcs[cur.words[0]];
            } else {
              a = FindArc(st[s], cur.words[i]);
            }
            if (!a) {  // "no A B, but have A B C": the line is not added
              found = false;
              break;
            }
            s = a->to;
            hist_state = s;
          }
        }
        if (!found) {
          hist_state = 0;
          continue;
        }
        const int32_t to = add_state();
        if (!grow(hist_state, cur.words[n - 1], cur.logprob, to)) return ASRD_ERR_IO;
        // back-off target: the state of the longest proper suffix that exists
        int32_t target = 0;
        for (size_t from = 1; from < n; ++from) {
          int32_t s = 0;
          size_t i = from;
          for (; i < n; ++i) {
            const LmArc *a = nullptr;
            if (i == from) {
              if (cur.words[i] < (int32_t)st[0].arcs.size()) a = &st[0].arcs[cur.words[i]];
            } else {
              a = FindArc(st[s], cur.words[i]);
            }
            if (!a) break;
            s = a->to;
          }
          if (i == n) {
            target = s;
            break;
          }
          if (from + 1 == n) return ASRD_ERR_IO;  // (the reference asserts: the last word is a unigram)
        }
        st[to].backoff_id = target;
        st[to].backoff_prob = cur.backoff;
      }
      prev = cur;
    }
  }
  // ---- the file (ArpaLm::Write arpa2fsa.h:441-480, Fsa::Write arpa2fsa.cc:9-68)
  FILE *fp = fopen(out_path, "wb");
  if (!fp) return ASRD_ERR_IO;
  bool okw = true;
  auto put = [&](const void *p, size_t bytes) { okw = okw && fwrite(p, 1, bytes, fp) == bytes; };
  const int32_t hdr[3] = {bos, eos, unk};
  put(hdr, sizeof(hdr));
  const uint64_t n_orders = order;
  put(&n_orders, sizeof(n_orders));
  put(declared.data(), 4 * declared.size());
  const int32_t n_states = (int32_t)st.size();
  put(&n_states, 4);
  int32_t n_arcs = 0;
  for (const LmState &s : st) {
    const int32_t an = (int32_t)s.arcs.size();
    put(&an, 4);
    put(&s.backoff_prob, 4);
    put(&s.backoff_id, 4);
    n_arcs += an;
  }
  put(&n_arcs, 4);
  for (const LmState &s : st)
    if (!s.arcs.empty()) put(s.arcs.data(), sizeof(LmArc) * s.arcs.size());
  okw = fclose(fp) == 0 && okw;
  return okw ? ASRD_OK : ASRD_ERR_IO;
}

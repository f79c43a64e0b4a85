// Reader of the reference's ".u" input format with its site-pattern compression, host C++ (no device code).
//
// What it reproduces (src/readata.cpp): the top lines (read_datafile_top_lines :891-1036), the locus header line
// (parse_locus_info :618-866), and per mutation model
//   I / J  findsegsites :34-178 + readseqIS :347-497: a column is kept iff it is segregating, has exactly two
//          states, and every base in it is one of acgt; kept columns are recoded 0 = the first gene's base, 1 = the other
//   H      readseqHKY :246-345: a c g t/u -> 0 1 2 3, n - . -> gap; base frequencies over all bases read; columns with a gap
//          dropped (eliminategaps :224-243); identical columns merged with multiplicities in order of first appearance
//          (sortseq :202-222)
//   S / J  readseqSW :500-588: allele lengths of the linked stepwise parts, their minimum and maximum
// The reference exits through IM_err on a malformed file; here every such case is a negative return code and a message.
#include "../../include/ima2p_b200.h"
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <string>
#include <vector>

extern "C" void ima2p_internal_set_error(const char *msg);

namespace {

constexpr int kNameLen = 10;          // LENGENENAME, imamp.hpp
constexpr int kMinStrLength = 3;      // MINSTRLENGTH, imamp.hpp:144
constexpr int kMaxPopsU = 10, kMaxLinkedU = 15;

struct ULocus {
  std::string name;
  int model = 0, numgenes = 0, numbases = 0, numsites = 0, totsites = 0, nlinked = 1, nAlinked = 0;
  double hval = 1.0;
  bool hval_given = false;            // an inheritance scalar stood on the header line (the reference echoes it differently)
  bool model_digit = false;           // the model letter carried a count (S2, J1): the reference names these SW_M / IS+SW_M
  std::vector<int> samppop;
  std::vector<int> seq;               // [numgenes][numsites]
  std::vector<int> mult;              // [numsites] (HKY)
  std::vector<int> A;                 // [nlinked][numgenes] (part 0 of a J locus is unused)
  std::vector<int> minA, maxA;        // [nlinked]
  std::vector<double> urate;          // mutation rates per year given on the header line
  double pi[4] = {0, 0, 0, 0};
};

int ufail(int code, const std::string &msg) {
  ima2p_internal_set_error(msg.c_str());
  return code;
}

std::vector<std::string> split_ws(const std::string &s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i < s.size()) {
    while (i < s.size() && isspace((unsigned char)s[i])) i++;
    size_t j = i;
    while (j < s.size() && !isspace((unsigned char)s[j])) j++;
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j;
  }
  return out;
}

bool is_acgt(char c) { return c == 'a' || c == 'c' || c == 'g' || c == 't'; }

}  // namespace

struct ima2p_dataset {
  int npops = 0;
  std::string tree, title;
  std::vector<std::string> popnames, comments;      // comments: the '#' lines under the title, without the '#'
  std::vector<ULocus> loci;
};

namespace {

// the bases of one data line: everything after the 10-character name (and, for J loci, after the allele numbers),
// blanks dropped; readseqIS skips ' ' only, readseqHKY any white space -- tabs are rejected for I/J below
int read_is_locus(ULocus &L, const std::vector<std::string> &lines, size_t first, int li) {
  const int n = L.numgenes, nb = L.numbases;
  std::vector<std::string> bases(n);
  if (L.model == IMA2P_MODEL_JOINT) {
    L.A.assign((size_t)L.nlinked * n, 0);
    L.minA.assign(L.nlinked, 10000); L.maxA.assign(L.nlinked, -1);
    L.minA[0] = L.maxA[0] = 0;
  }
  for (int i = 0; i < n; i++) {
    const std::string &ln = lines[first + i];
    if ((int)ln.size() < kNameLen) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ", data line " + std::to_string(i) + ": no gene name");
    size_t pos = kNameLen;
    if (L.model == IMA2P_MODEL_JOINT) {
      for (int ai = 1; ai < L.nlinked; ai++) {
        while (pos < ln.size() && isspace((unsigned char)ln[pos])) pos++;
        char *end = nullptr;
        const long a = strtol(ln.c_str() + pos, &end, 10);
        if (end == ln.c_str() + pos) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ", data line " + std::to_string(i) + ": missing str data");
        if (a == 0) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": null alleles not allowed in STR data");
        if (a <= kMinStrLength) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": STR repeat numbers must exceed 3");
        pos = end - ln.c_str();
        L.A[(size_t)ai * n + i] = (int)a;
        if (a > L.maxA[ai]) L.maxA[ai] = (int)a;
        if (a < L.minA[ai]) L.minA[ai] = (int)a;
      }
    }
    std::string &b = bases[i];
    for (; pos < ln.size(); pos++) {
      const char c = (char)tolower((unsigned char)ln[pos]);
      if (c == ' ' || c == '\r') continue;
      if (isdigit((unsigned char)c)) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": formatting of input file causes wrong lines to be read as data");
      b.push_back(c);
    }
    if ((int)b.size() < nb) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ", data line " + std::to_string(i) + ": sequence length shorter than expected");
    if ((int)b.size() > nb) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ", data line " + std::to_string(i) + ": characters extend past the stated sequence length");
  }
  // findsegsites :34-178
  std::vector<char> zeroc(nb), altc(nb, ' '), bad(nb, 0), seg(nb, 0);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < nb; j++) {
      const char c = bases[i][j];
      if (i == 0) { zeroc[j] = c; if (!is_acgt(c)) bad[j] = 1; continue; }
      if (!is_acgt(c)) bad[j] = 1;
      if (!bad[j] && c != zeroc[j]) {
        if (altc[j] == ' ') altc[j] = c;
        else if (c != altc[j]) bad[j] = 1;
        seg[j] = 1;
      }
    }
  L.numsites = 0;
  for (int j = 0; j < nb; j++) { if (bad[j]) seg[j] = 0; L.numsites += seg[j]; }
  // readseqIS :447-472
  L.seq.assign((size_t)n * L.numsites, 0);
  for (int i = 0; i < n; i++) {
    int s = 0;
    for (int j = 0; j < nb; j++)
      if (seg[j]) { L.seq[(size_t)i * L.numsites + s] = (i == 0 || bases[i][j] == bases[0][j]) ? 0 : 1; s++; }
  }
  return 0;
}

int read_hky_locus(ULocus &L, const std::vector<std::string> &lines, size_t first, size_t *used, int li) {
  const int n = L.numgenes;
  int ns = L.numbases;
  std::vector<int> seq((size_t)n * ns, 0);
  // interleaved blocks of numgenes lines until every site has been read (:259-329)
  int k = 0;
  size_t at = first;
  while (k < ns) {
    int kend = k;
    for (int i = 0; i < n; i++, at++) {
      if (at >= lines.size()) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": file ends inside HKY data");
      const std::string &ln = lines[at];
      if ((int)ln.size() < kNameLen) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": no gene name");
      int j = k;
      for (size_t pos = kNameLen; pos < ln.size(); pos++) {
        const char c = ln[pos];
        if (isspace((unsigned char)c)) continue;
        if (j >= ns) return ufail(IMA2P_E_ARG, "HKY data problem locus " + std::to_string(li) + " gene# " + std::to_string(i) + " site# " + std::to_string(j));
        int v;
        switch (c) {
          case 'a': case 'A': v = 0; break;
          case 'c': case 'C': v = 1; break;
          case 'g': case 'G': v = 2; break;
          case 't': case 'T': case 'u': case 'U': v = 3; break;
          case 'n': case 'N': case '-': case '.': v = -1; break;
          default: return ufail(IMA2P_E_ARG, "BAD BASE in locus " + std::to_string(li) + " species " + std::to_string(i + 1) + " base " + std::to_string(j + 1));
        }
        if (v >= 0) L.pi[v] += 1.0;
        seq[(size_t)i * ns + j] = v;
        j++;
      }
      if (i == 0) kend = j;
      else if (j != kend) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": HKY data lines of unequal length");
    }
    if (kend == k) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": empty HKY data block");
    k = kend;
  }
  *used = at - first;
  // eliminategaps :224-243: every column that holds a gap goes
  std::vector<int> keep;
  for (int j = 0; j < ns; j++) {
    bool gap = false;
    for (int i = 0; i < n; i++) gap |= seq[(size_t)i * ns + j] == -1;
    if (!gap) keep.push_back(j);
  }
  L.totsites = (int)keep.size();
  // sortseq :202-222: identical columns are merged into the first of them
  std::vector<int> pat, mult;
  for (int j : keep) {
    int hit = -1;
    for (size_t q = 0; q < pat.size() && hit < 0; q++) {
      bool same = true;
      for (int i = 0; i < n && same; i++) same = seq[(size_t)i * ns + j] == seq[(size_t)i * ns + pat[q]];
      if (same) hit = (int)q;
    }
    if (hit < 0) { pat.push_back(j); mult.push_back(1); } else mult[hit]++;
  }
  L.numsites = (int)pat.size();
  L.mult = mult;
  L.seq.assign((size_t)n * L.numsites, 0);
  for (int i = 0; i < n; i++)
    for (int q = 0; q < L.numsites; q++) L.seq[(size_t)i * L.numsites + q] = seq[(size_t)i * ns + pat[q]];
  double tot = L.pi[0] + L.pi[1] + L.pi[2] + L.pi[3];
  if (!(tot > 0)) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": no base left after the columns with gaps were removed");
  for (int b = 0; b < 4; b++) L.pi[b] = L.pi[b] / tot;
  return 0;
}

int read_sw_locus(ULocus &L, const std::vector<std::string> &lines, size_t first, int li) {
  const int n = L.numgenes;
  L.numsites = 0;
  L.A.assign((size_t)L.nlinked * n, 0);
  L.minA.assign(L.nlinked, 10000); L.maxA.assign(L.nlinked, -1);
  for (int i = 0; i < n; i++) {
    const std::string &ln = lines[first + i];
    if ((int)ln.size() < kNameLen) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": no gene name");
    size_t pos = kNameLen;
    for (int ai = 0; ai < L.nAlinked; ai++) {
      while (pos < ln.size() && isspace((unsigned char)ln[pos])) pos++;
      char *end = nullptr;
      const long a = strtol(ln.c_str() + pos, &end, 10);
      if (end == ln.c_str() + pos) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ", data line " + std::to_string(i) + ": missing str data");
      if (a == 0) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": null alleles not allowed in STR data");
      if (a <= kMinStrLength) return ufail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": STR repeat numbers must exceed 3");
      pos = end - ln.c_str();
      L.A[(size_t)ai * n + i] = (int)a;
      if (a > L.maxA[ai]) L.maxA[ai] = (int)a;
      if (a < L.minA[ai]) L.minA[ai] = (int)a;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

static int dataset_read_impl(const char *path, ima2p_dataset **out);
// no exception crosses the C boundary: a malformed file is an error code, whatever it breaks on the way
int ima2p_dataset_read(const char *path, ima2p_dataset **out) {
  try {
    return dataset_read_impl(path, out);
  } catch (const std::exception &ex) {
    return ufail(IMA2P_E_ARG, std::string("malformed data file: ") + ex.what());
  } catch (...) {
    return ufail(IMA2P_E_ARG, "malformed data file");
  }
}
static int dataset_read_impl(const char *path, ima2p_dataset **out) {
  if (!path || !out) return ufail(IMA2P_E_ARG, "dataset_read: bad argument");
  FILE *f = fopen(path, "r");
  if (!f) return ufail(IMA2P_E_ARG, std::string("data file not found or can't be opened: ") + path);
  std::vector<std::string> lines;
  {
    std::string cur;
    int ch;
    while ((ch = fgetc(f)) != EOF) {
      if (ch == '\n') { lines.push_back(cur); cur.clear(); } else cur.push_back((char)ch);
    }
    if (!cur.empty()) lines.push_back(cur);
    fclose(f);
  }
  for (auto &l : lines) while (!l.empty() && l.back() == '\r') l.pop_back();
  ima2p_dataset *D = new ima2p_dataset();
  auto bail = [&](int code, const std::string &m) { delete D; return ufail(code, m); };
  size_t at = 0;
  if (lines.empty()) return bail(IMA2P_E_ARG, "empty data file");
  D->title = lines[at++];
  while (at < lines.size() && !lines[at].empty() && lines[at][0] == '#') D->comments.push_back(lines[at++].substr(1));
  // npops, population names, tree string (optional for two populations), number of loci: a token stream (:929-1020)
  std::vector<std::string> tok;
  auto need = [&](size_t k) { while (tok.size() < k && at < lines.size()) { auto t = split_ws(lines[at++]); tok.insert(tok.end(), t.begin(), t.end()); } return tok.size() >= k; };
  if (!need(1)) return bail(IMA2P_E_ARG, "data file ends before the number of populations");
  D->npops = atoi(tok[0].c_str());
  if (D->npops < 1 || D->npops > kMaxPopsU) return bail(IMA2P_E_ARG, "number of populations must be between 1 and 10");
  if (!need(2 + (size_t)D->npops)) return bail(IMA2P_E_ARG, "data file ends before the population tree");
  for (int i = 0; i < D->npops; i++) D->popnames.push_back(tok[1 + i]);
  int nloci = 0;
  const std::string ts = tok[1 + D->npops];
  if (ts.size() >= 5 || ts.size() == 1) {
    D->tree = ts;
    if (!need(3 + (size_t)D->npops)) return bail(IMA2P_E_ARG, "data file ends before the number of loci");
    nloci = atoi(tok[2 + D->npops].c_str());
    if (tok.size() != 3 + (size_t)D->npops) return bail(IMA2P_E_ARG, "unexpected text after the number of loci");
  } else {
    if (D->npops != 2) return bail(IMA2P_E_ARG, "no population string given in file, but # of populations is greater than 2");
    nloci = atoi(ts.c_str());
    D->tree = "(0,1):2";
    if (tok.size() != 2 + (size_t)D->npops) return bail(IMA2P_E_ARG, "unexpected text after the number of loci");
  }
  if (nloci < 1 || nloci > 1000) return bail(IMA2P_E_ARG, "number of loci must be between 1 and MAXLOCI (1000)");
  D->loci.resize(nloci);
  for (int li = 0; li < nloci; li++) {
    ULocus &L = D->loci[li];
    if (at >= lines.size()) return bail(IMA2P_E_ARG, "data file ends before locus " + std::to_string(li));
    const std::vector<std::string> h = split_ws(lines[at++]);
    if ((int)h.size() < D->npops + 3) return bail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": header line too short");
    L.name = h[0];
    for (int i = 0; i < D->npops; i++) {
      L.samppop.push_back(atoi(h[1 + i].c_str()));
      if (L.samppop.back() < 0) return bail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": negative sample size");
      L.numgenes += L.samppop.back();
    }
    L.numbases = atoi(h[1 + D->npops].c_str());
    if (L.numbases < 0) return bail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": negative sequence length");
    const std::string &mt = h[2 + D->npops];
    const int digits = mt.size() > 1 && isdigit((unsigned char)mt[1]) ? atoi(mt.c_str() + 1) : 0;
    L.model_digit = mt.size() > 1 && isdigit((unsigned char)mt[1]);
    switch (toupper((unsigned char)mt[0])) {
      case 'H': L.model = IMA2P_MODEL_HKY; L.nlinked = 1; break;
      case 'S': L.model = IMA2P_MODEL_SW; L.nAlinked = digits ? digits : 1; L.nlinked = L.nAlinked; break;
      case 'J': L.model = IMA2P_MODEL_JOINT; L.nAlinked = digits ? digits : 1; L.nlinked = L.nAlinked + 1; break;
      default: L.model = IMA2P_MODEL_IS; L.nlinked = 1; break;        // 'I' and, as in the reference, anything else
    }
    if (L.nlinked < 1 || L.nlinked > kMaxLinkedU) return bail(IMA2P_E_ARG, "the number of linked parts is less than 1 or greater than 15");
    if (L.numgenes < 2) return bail(IMA2P_E_ARG, "locus " + std::to_string(li) + ": fewer than two genes");
    size_t q = 3 + D->npops;
    if (q < h.size() && h[q][0] == 'A') return bail(IMA2P_E_UNSUPPORTED, "genes of unknown origin (assignment model) are not supported");
    if (q < h.size()) { L.hval = atof(h[q].c_str()); L.hval_given = true; q++; } else L.hval = 1.0;
    for (; q < h.size() && (int)L.urate.size() < L.nlinked; q++) {
      if (h[q][0] == 'A') return bail(IMA2P_E_UNSUPPORTED, "genes of unknown origin (assignment model) are not supported");
      if (h[q][0] == '(') continue;             // a prior range on the rate: used only with -p options outside this path
      L.urate.push_back(atof(h[q].c_str()));
    }
    if (at + L.numgenes > lines.size()) return bail(IMA2P_E_ARG, "data file ends inside locus " + std::to_string(li));
    int rc = 0;
    size_t used = L.numgenes;
    if (L.model == IMA2P_MODEL_IS || L.model == IMA2P_MODEL_JOINT) {
      if (L.numbases > 0) rc = read_is_locus(L, lines, at, li);
      else if (L.model == IMA2P_MODEL_JOINT) rc = ufail(IMA2P_E_ARG, "joint locus without sequence");
    } else if (L.model == IMA2P_MODEL_HKY) rc = read_hky_locus(L, lines, at, &used, li);
    else rc = read_sw_locus(L, lines, at, li);
    if (rc) { delete D; return rc; }
    at += used;
  }
  *out = D;
  return IMA2P_OK;
}

void ima2p_dataset_free(ima2p_dataset *d) { delete d; }

int ima2p_dataset_dims(const ima2p_dataset *d, int *npops, int *nloci, char *tree, int tree_len) {
  if (!d) return ufail(IMA2P_E_ARG, "null dataset");
  if (npops) *npops = d->npops;
  if (nloci) *nloci = (int)d->loci.size();
  if (tree && tree_len > 0) { strncpy(tree, d->tree.c_str(), tree_len - 1); tree[tree_len - 1] = '\0'; }
  return IMA2P_OK;
}

// the text the reference echoes into its report (readata.cpp:916-963): kind 0 = title line (index 0) and the '#' lines
// under it (index 1..), kind 1 = population names; returns IMA2P_E_ARG past the last one
int ima2p_dataset_text(const ima2p_dataset *d, int kind, int index, char *buf, int buf_len) {
  if (!d || !buf || buf_len < 1 || index < 0) return ufail(IMA2P_E_ARG, "dataset_text: bad argument");
  const std::string *s = nullptr;
  if (kind == 0) s = index == 0 ? &d->title : index <= (int)d->comments.size() ? &d->comments[index - 1] : nullptr;
  else if (kind == 1) s = index < (int)d->popnames.size() ? &d->popnames[index] : nullptr;
  if (!s) return ufail(IMA2P_E_ARG, "dataset_text: no such line");
  strncpy(buf, s->c_str(), buf_len - 1); buf[buf_len - 1] = '\0';
  return IMA2P_OK;
}

// info[8] = model, numgenes, numsites, totsites, numbases, nlinked, number of mutation rates on the header line,
//           bit 0: the model letter carried a count (S2, J1; readata.cpp:664-692), bit 1: an inheritance scalar was given (:729-734, 826-831)
int ima2p_dataset_locus(const ima2p_dataset *d, int li, int *info, double *hval, int *samppop, char *name, int name_len) {
  if (!d || li < 0 || li >= (int)d->loci.size() || !info) return ufail(IMA2P_E_ARG, "dataset_locus: bad argument");
  const ULocus &L = d->loci[li];
  info[0] = L.model; info[1] = L.numgenes; info[2] = L.numsites; info[3] = L.totsites; info[4] = L.numbases; info[5] = L.nlinked;
  info[6] = (int)L.urate.size(); info[7] = (L.model_digit ? 1 : 0) | (L.hval_given ? 2 : 0);
  if (hval) *hval = L.hval;
  if (samppop) for (int i = 0; i < d->npops; i++) samppop[i] = L.samppop[i];
  if (name && name_len > 0) { strncpy(name, L.name.c_str(), name_len - 1); name[name_len - 1] = '\0'; }
  return IMA2P_OK;
}

// seq[numgenes][numsites], mult[numsites] (HKY), A[nlinked][numgenes], minA/maxA[nlinked], pi[4] (HKY), urate[info[6]];
// any pointer may be NULL
int ima2p_dataset_locus_data(const ima2p_dataset *d, int li, int *seq, int *mult, int *A, int *minA, int *maxA, double *pi, double *urate) {
  if (!d || li < 0 || li >= (int)d->loci.size()) return ufail(IMA2P_E_ARG, "dataset_locus_data: bad argument");
  const ULocus &L = d->loci[li];
  if (seq) for (size_t i = 0; i < L.seq.size(); i++) seq[i] = L.seq[i];
  if (mult) for (size_t i = 0; i < L.mult.size(); i++) mult[i] = L.mult[i];
  if (A) for (size_t i = 0; i < L.A.size(); i++) A[i] = L.A[i];
  if (minA) for (size_t i = 0; i < L.minA.size(); i++) minA[i] = L.minA[i];
  if (maxA) for (size_t i = 0; i < L.maxA.size(); i++) maxA[i] = L.maxA[i];
  if (pi) for (int b = 0; b < 4; b++) pi[b] = L.pi[b];
  if (urate) for (size_t i = 0; i < L.urate.size(); i++) urate[i] = L.urate[i];
  return IMA2P_OK;
}

// ---- the .ti file of sampled genealogies (written in M mode, read back in L mode) ---------------------------------
// header block + "VALUESSTART" (ima_main_mpi.cpp:2123-2141), then one line per sampled genealogy, every value as
// "%.6f\t" (savegenealogyfile, output.cpp:662-685); read back by loadgenealogyvalues (ima_main_mpi.cpp:3216-3440)
int ima2p_ti_create(const char *path, const char *header_text) {
  if (!path) return ufail(IMA2P_E_ARG, "ti_create: bad argument");
  FILE *f = fopen(path, "w");
  if (!f) return ufail(IMA2P_E_ARG, "Error creating file for holding genealogy information");
  const char *bar = "-------------------------------------------\n\n";
  fprintf(f, "%s", bar);
  fprintf(f, "Header for genealogy file:  %s\n\n", path);
  fprintf(f, "%s", bar);
  fprintf(f, "%s\n", header_text ? header_text : "");
  fprintf(f, "%s", bar);
  fprintf(f, "End of header for genealogy file:  %s\n\n", path);
  fprintf(f, "%s", bar);
  fprintf(f, "VALUESSTART\n");
  fclose(f);
  return IMA2P_OK;
}

int ima2p_ti_append(const char *path, const float *rows, long long nrows, int rowlen) {
  if (!path || !rows || nrows < 0 || rowlen < 1) return ufail(IMA2P_E_ARG, "ti_append: bad argument");
  FILE *f = fopen(path, "a");
  if (!f) return ufail(IMA2P_E_ARG, "Error opening treeinfosave file for writing");
  for (long long j = 0; j < nrows; j++) {
    for (int i = 0; i < rowlen; i++) fprintf(f, "%.6f\t", (float)rows[(size_t)j * rowlen + i]);
    fprintf(f, "\n");
  }
  fclose(f);
  return IMA2P_OK;
}

// rows == NULL: only counts the genealogies in the file.  Otherwise loads up to max_rows of them (all when the file
// holds fewer); a line with fewer or more than rowlen values is an error, as in the reference (IMERR_TIFILE).
int ima2p_ti_load(const char *path, int rowlen, float *rows, long long max_rows, long long *nrows_out) {
  if (!path || rowlen < 1 || !nrows_out) return ufail(IMA2P_E_ARG, "ti_load: bad argument");
  FILE *f = fopen(path, "r");
  if (!f) return ufail(IMA2P_E_ARG, " cannot open .ti file");
  std::string line;
  int ch;
  bool started = false;
  long long n = 0;
  auto getline = [&]() { line.clear(); while ((ch = fgetc(f)) != EOF && ch != '\n') line.push_back((char)ch); return ch != EOF || !line.empty(); };
  while (getline()) if (line.find("VALUESSTART") != std::string::npos) { started = true; break; }
  if (!started) { fclose(f); return ufail(IMA2P_E_ARG, "no VALUESSTART line in .ti file"); }
  while (getline()) {
    if (rows && n >= max_rows) break;
    const char *c = line.c_str();
    int got = 0;
    for (;;) {
      while (*c && isspace((unsigned char)*c)) c++;
      if (!*c) break;
      char *end = nullptr;
      const float v = strtof(c, &end);
      if (end == c) { fclose(f); return ufail(IMA2P_E_ARG, "Problem in .ti file: not a number"); }
      if (got < rowlen && rows) rows[(size_t)n * rowlen + got] = v;
      got++;
      c = end;
    }
    if (got == 0) continue;
    if (got < rowlen) { fclose(f); return ufail(IMA2P_E_ARG, "Problem in .ti file, too few values per genealogy, .ti file may have been generated with a different program"); }
    if (got > rowlen) { fclose(f); return ufail(IMA2P_E_ARG, "Problem in .ti file, too many values per genealogy, .ti file may have been generated with a different program"); }
    n++;
  }
  fclose(f);
  *nrows_out = n;
  return IMA2P_OK;
}

}  // extern "C"

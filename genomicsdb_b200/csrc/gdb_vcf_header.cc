// See gdb_vcf_header.h.  Plain host C++ (no htslib): the header arithmetic htslib does for the reference, restated.
#include "gdb_vcf_header.h"

#include <fcntl.h>
#include <unistd.h>
#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <climits>
#include <cstring>
#include <strings.h>
#include <fstream>
#include <sstream>
#include <stdexcept>

namespace genomicsdb_b200 {

namespace {
bool is_escaped(const char* min, const char* str) {   // an odd number of backslashes before str
  int n = 0;
  while (--str >= min && *str == '\\') n++;
  return n % 2;
}
}  // namespace

VCFHeaderLines::VCFHeaderLines(bool write_mode) {
  if (write_mode) {
    append("##fileformat=VCFv4.2");
    append("##FILTER=<ID=PASS,Description=\"All filters passed\">");   // the filter PASS must appear first in the dictionary
  }
}

// vcf.c:289-391.  *len: characters consumed (0: not a ## line, -1: malformed key)
std::unique_ptr<VCFHeaderRec> VCFHeaderLines::parse_line(const char* line, int* len) {
  const char* p = line;
  if (p[0] != '#' || p[1] != '#') { *len = 0; return nullptr; }
  p += 2;
  const char* q = p;
  while (*q && *q != '=' && *q != '\n') q++;
  const int n = (int)(q - p);
  if (*q != '=' || !n) { *len = -1; return nullptr; }
  std::unique_ptr<VCFHeaderRec> rec(new VCFHeaderRec);
  rec->key.assign(p, n);
  p = ++q;
  if (*p != '<') {   // generic field, e.g. ##samtoolsVersion=0.1.18-r579
    while (*q && *q != '\n') q++;
    rec->value.assign(p, q - p);
    *len = (int)(q - line) + (*q ? 1 : 0);
    return rec;
  }
  rec->structured = true;
  int nopen = 1;
  while (*q && *q != '\n' && nopen > 0) {
    p = ++q;
    while (*q && *q == ' ') { p++; q++; }
    if (p == q && *q && (isalpha((unsigned char)*q) || *q == '_')) {   // ^[A-Za-z_][0-9A-Za-z_.]*$
      q++;
      while (*q && (isalnum((unsigned char)*q) || *q == '_' || *q == '.')) q++;
    }
    const int nk = (int)(q - p);
    int m = 0;
    while (*q && *q == ' ') { q++; m++; }
    if (*q != '=' || !nk) {   // wrong format: the line is dropped
      while (*q && *q != '\n') q++;
      *len = (int)(q - line) + (*q ? 1 : 0);
      return nullptr;
    }
    std::string key(p, q - p - m);
    p = ++q;
    while (*q && *q == ' ') { p++; q++; }
    const bool quoted = *p == '"';
    if (quoted) { p++; q++; }
    while (*q && *q != '\n') {
      if (quoted) { if (*q == '"' && !is_escaped(p, q)) break; }
      else {
        if (*q == '<') nopen++;
        if (*q == '>') nopen--;
        if (!nopen) break;
        if (*q == ',' && nopen == 1) break;
      }
      q++;
    }
    const char* r = q;
    while (r > p && r[-1] == ' ') r--;
    std::string val(p, r - p);
    if (quoted) val = "\"" + val + "\"";
    rec->kv.emplace_back(std::move(key), std::move(val));
    if (quoted && *q == '"') q++;
    if (*q == '>') { nopen--; q++; }
  }
  while (*q && *q != '\n') q++;   // trailing junk is dropped
  *len = (int)(q - line) + (*q ? 1 : 0);
  return rec;
}

// vcf.c:407-545: true when the line enters a dictionary (or is a structured line of another kind)
bool VCFHeaderLines::register_rec(VCFHeaderRec* r) {
  auto parse_idx = [](const std::string& v, int& idx) -> bool {
    char* end = nullptr;
    long x = strtol(v.c_str(), &end, 10);
    if (*end || x < 0 || x >= INT_MAX - 1) return false;
    idx = (int)x;
    return true;
  };
  auto take_idx = [](std::vector<bool>& used, int& n, int& id, const std::string& tag) {   // bcf_hdr_set_idx
    if (id == -1) id = n++;
    else if (id < n && id < (int)used.size() && used[id])   // the reference's htslib ends the process here (vcf.c:389-394)
      throw std::runtime_error("Conflicting IDX=" + std::to_string(id) + " lines in the header dictionary, the new tag is " + tag);
    if (id >= n) n = id + 1;
    if ((int)used.size() < n) used.resize(n, false);
    used[id] = true;
  };
  if (r->key == "contig") {
    r->type = HL_CTG;
    int i = r->find_key("length");
    if (i >= 0) { int len; if (sscanf(r->kv[i].second.c_str(), "%d", &len) != 1) return false; }   // a length that is no number: the line is dropped
    i = r->find_key("ID");
    if (i < 0) return false;
    const std::string id = r->kv[i].second;
    if (ctgs_.count(id)) return false;   // already present
    int idx = -1;
    const int k = r->find_key("IDX");
    if (k >= 0 && !parse_idx(r->kv[k].second, idx)) return false;
    CtgInfo& c = ctgs_[id];
    c.id = idx; c.rec = r;
    take_idx(ctg_used_, n_ctgs_, c.id, id);
    if (idx == -1) r->kv.emplace_back("IDX", std::to_string(c.id));
    return true;
  }
  if (r->key == "INFO") r->type = HL_INFO;
  else if (r->key == "FILTER") r->type = HL_FLT;
  else if (r->key == "FORMAT") r->type = HL_FMT;
  else if (!r->kv.empty()) { r->type = HL_STR; return true; }
  else return false;
  const std::string* id = nullptr;
  uint32_t type = UINT32_MAX, var = UINT32_MAX;
  int num = -1, idx = -1;
  for (auto& kv : r->kv) {
    if (kv.first == "ID") id = &kv.second;
    else if (kv.first == "IDX") { if (!parse_idx(kv.second, idx)) return false; }
    else if (kv.first == "Type") {
      if (kv.second == "Integer") type = HT_INT;
      else if (kv.second == "Float") type = HT_REAL;
      else if (kv.second == "Flag") type = HT_FLAG;
      else type = HT_STR;   // String, Character, anything else
    } else if (kv.first == "Number") {
      if (kv.second == "A") var = VL_A;
      else if (kv.second == "R") var = VL_R;
      else if (kv.second == "G") var = VL_G;
      else if (kv.second == ".") var = VL_VAR;
      else { sscanf(kv.second.c_str(), "%d", &num); var = VL_FIXED; }
      if (var != VL_FIXED) num = 0xfffff;
    }
  }
  if (r->type == HL_INFO || r->type == HL_FMT) {
    if (type == UINT32_MAX) type = HT_STR;
    if (var == UINT32_MAX) var = VL_VAR;
  }
  const uint32_t info = (((uint32_t)num) & 0xfffff) << 12 | (var & 0xf) << 8 | (type & 0xf) << 4 | ((uint32_t)r->type & 0xf);
  if (!id) return false;
  auto it = ids_.find(*id);
  if (it != ids_.end()) {   // the ID is known: fine as long as no line of this kind carries it
    IdInfo& e = it->second;
    if (e.rec[r->type]) return false;
    e.info[r->type] = info; e.rec[r->type] = r;
    if (idx == -1) r->kv.emplace_back("IDX", std::to_string(e.id));
    return true;
  }
  IdInfo& e = ids_[*id];
  e.info[r->type] = info; e.rec[r->type] = r; e.id = idx;
  take_idx(id_used_, n_ids_, e.id, *id);
  if (idx == -1) r->kv.emplace_back("IDX", std::to_string(e.id));
  return true;
}

// vcf.c:547-588
int VCFHeaderLines::add_rec(std::unique_ptr<VCFHeaderRec> r) {
  if (!r) return 0;
  r->type = HL_GEN;
  if (!register_rec(r.get())) {
    if (r->type != HL_GEN) return 0;   // one of the hashed kinds: already present
    for (auto& o : recs_) {
      if (o->type != HL_GEN || o->key != r->key) continue;
      if (r->key == "fileformat" || o->value == r->value) return 0;
    }
  }
  const int ret = r->type == HL_GEN ? 0 : 1;
  recs_.push_back(std::move(r));
  return ret;
}

bool VCFHeaderLines::append(const std::string& line) {
  int len = 0;
  auto r = parse_line(line.c_str(), &len);
  if (!r) return false;
  add_rec(std::move(r));
  return true;
}

// vcf.c:645-718
bool VCFHeaderLines::parse(const std::string& text, bool sample_line_required) {
  const char* p = text.c_str();
  int len = 0;
  // "fileformat" must come first: the reference's htslib still reads the rest but reports failure (vcf.c:652-658,714-720), and
  // the reader above it then has no header
  bool ok = true;
  {
    auto first = parse_line(p, &len);
    if (!first || strcasecmp(first->key.c_str(), "fileformat") != 0) ok = false;
    add_rec(std::move(first));
  }
  add_rec(parse_line("##FILTER=<ID=PASS,Description=\"All filters passed\">", &len));   // PASS first in the dictionary
  int done = 0;
  do {
    for (;;) {
      auto r = parse_line(p, &len);
      if (!r) {
        if (len < 0) return false;
        if (len > 0) { p += len; continue; }   // a malformed structured line was dropped
        break;
      }
      add_rec(std::move(r));
      p += len;
    }
    if (strncmp("#CHROM\tPOS", p, 10) != 0) {
      const char* eol = strchr(p, '\n');
      if (eol) p = eol + 1; else done = -1;
    } else done = 1;
  } while (!done);
  if (done < 0) return ok && !sample_line_required;
  if (!ok) return false;
  // sample line: columns past the ninth
  int i = 0;
  const char* q;
  for (p = q = p;; ++q) {
    if (*q != '\t' && *q != 0 && *q != '\n') continue;
    if (++i > 9 && !add_sample(std::string(p, q - p))) ok = false;   // a sample named twice: every name is still read, the parse fails (vcf.c:122-140)
    if (*q == 0 || *q == '\n') break;
    p = q + 1;
  }
  return ok;
}

// vcf.c:736-795 (the branch with a key)
void VCFHeaderLines::remove(int type, const std::string& id) {
  VCFHeaderRec* victim = nullptr;
  if (type == HL_FLT || type == HL_INFO || type == HL_FMT) {
    auto it = ids_.find(id);
    if (it == ids_.end() || !it->second.rec[type]) return;
    victim = it->second.rec[type];
    it->second.rec[type] = nullptr;   // the dictionary entry (and its index) stays
  } else if (type == HL_CTG) {
    auto it = ctgs_.find(id);
    if (it == ctgs_.end() || !it->second.rec) return;
    victim = it->second.rec; it->second.rec = nullptr;
  } else {
    for (auto& r : recs_) {
      if (r->type != type) continue;
      if (type == HL_GEN) { if (r->key == id) { victim = r.get(); break; } }
      else { int j = r->find_key("ID"); if (j >= 0 && r->kv[j].second == id) { victim = r.get(); break; } }
    }
  }
  if (!victim) return;
  for (size_t i = 0; i < recs_.size(); ++i) if (recs_[i].get() == victim) { recs_.erase(recs_.begin() + i); break; }
}

int VCFHeaderLines::id2int(const std::string& id) const { auto it = ids_.find(id); return it == ids_.end() ? -1 : it->second.id; }
bool VCFHeaderLines::exists(int type, const std::string& id) const {
  auto it = ids_.find(id);
  return it != ids_.end() && (it->second.info[type] & 0xf) != 0xf && it->second.rec[type] != nullptr;
}
const VCFHeaderRec* VCFHeaderLines::hrec(int type, const std::string& id) const {
  if (type == HL_CTG) { auto it = ctgs_.find(id); return it == ctgs_.end() ? nullptr : it->second.rec; }
  auto it = ids_.find(id);
  return it == ids_.end() ? nullptr : it->second.rec[type];
}
int VCFHeaderLines::id2type(int type, const std::string& id) const { auto it = ids_.find(id); return it == ids_.end() ? -1 : (int)((it->second.info[type] >> 4) & 0xf); }
int VCFHeaderLines::id2length(int type, const std::string& id) const { auto it = ids_.find(id); return it == ids_.end() ? -1 : (int)((it->second.info[type] >> 8) & 0xf); }
int VCFHeaderLines::id2number(int type, const std::string& id) const { auto it = ids_.find(id); return it == ids_.end() ? -1 : (int)(it->second.info[type] >> 12); }
int VCFHeaderLines::name2id(const std::string& contig) const { auto it = ctgs_.find(contig); return it == ctgs_.end() ? -1 : it->second.id; }

bool VCFHeaderLines::add_sample(const std::string& name) {
  if (std::find(samples_.begin(), samples_.end(), name) != samples_.end()) return false;
  samples_.push_back(name);
  return true;
}

// vcf.c:1837-1881
std::string VCFHeaderLines::format(bool is_bcf) const {
  std::string s;
  for (auto& r : recs_) {
    if (r->structured) {
      s += "##" + r->key + "=<";
      int nout = 0;
      for (auto& kv : r->kv) {
        if (!is_bcf && kv.first == "IDX") continue;
        if (nout) s += ',';
        s += kv.first + "=" + kv.second;
        nout++;
      }
      s += ">\n";
    } else s += "##" + r->key + "=" + r->value + "\n";
  }
  s += "#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO";
  if (!samples_.empty()) {
    s += "\tFORMAT";
    for (auto& n : samples_) s += "\t" + n;
  }
  s += "\n";
  return s;
}

std::vector<std::pair<std::string, int>> VCFHeaderLines::ids() const {
  std::vector<std::pair<std::string, int>> v;
  for (auto& kv : ids_) v.emplace_back(kv.first, kv.second.id);
  std::sort(v.begin(), v.end(), [](const std::pair<std::string, int>& a, const std::pair<std::string, int>& b) { return a.second < b.second; });
  return v;
}
std::vector<std::pair<std::string, int>> VCFHeaderLines::contigs() const {
  std::vector<std::pair<std::string, int>> v;
  for (auto& kv : ctgs_) v.emplace_back(kv.first, kv.second.id);
  std::sort(v.begin(), v.end(), [](const std::pair<std::string, int>& a, const std::pair<std::string, int>& b) { return a.second < b.second; });
  return v;
}

// ------------------------------------------------------------------------------------------------ IndexedFasta
IndexedFasta::~IndexedFasta() { if (fd_ >= 0) close(fd_); }

void IndexedFasta::open(const std::string& path) {
  std::ifstream fai(path + ".fai");
  if (!fai) throw std::runtime_error("cannot open the FASTA index " + path + ".fai");
  std::string line;
  while (std::getline(fai, line)) {
    std::istringstream is(line);
    std::string name; Entry e{};
    if (!(is >> name >> e.length >> e.offset >> e.line_bases >> e.line_width)) continue;
    index_[name] = e;
  }
  fd_ = ::open(path.c_str(), O_RDONLY);
  if (fd_ < 0) throw std::runtime_error("cannot open the reference genome " + path);
  unsigned char magic[4] = {0, 0, 0, 0};
  if (pread(fd_, magic, 4, 0) == 4 && magic[0] == 0x1f && magic[1] == 0x8b) {   // bgzip: blocks located through the .gzi index
    bgzf_ = true;
    std::ifstream gzi(path + ".gzi", std::ios::binary);
    if (!gzi) throw std::runtime_error("compressed reference genome without a .gzi index: " + path);
    uint64_t n = 0;
    gzi.read(reinterpret_cast<char*>(&n), 8);
    gzi_.assign(1, {0, 0});
    for (uint64_t i = 0; i < n; ++i) { uint64_t c = 0, u = 0; gzi.read(reinterpret_cast<char*>(&c), 8); gzi.read(reinterpret_cast<char*>(&u), 8); gzi_.emplace_back(c, u); }
  }
}

void IndexedFasta::read_raw(int64_t uoffset, char* dst, size_t n) {
  if (!bgzf_) {
    size_t got = 0;
    while (got < n) { ssize_t r = pread(fd_, dst + got, n - got, uoffset + (int64_t)got); if (r <= 0) break; got += (size_t)r; }
    if (got < n) memset(dst + got, 0, n - got);
    return;
  }
  size_t done = 0;
  while (done < n) {
    const int64_t u = uoffset + (int64_t)done;
    if (!(block_uoff_ >= 0 && u >= block_uoff_ && u < block_uoff_ + (int64_t)block_.size())) {
      // the block holding u: last index entry at or below it
      size_t lo = 0, hi = gzi_.size();
      while (hi - lo > 1) { size_t mid = (lo + hi) / 2; if ((int64_t)gzi_[mid].second <= u) lo = mid; else hi = mid; }
      uint8_t hdr[18];
      if (pread(fd_, hdr, 18, (off_t)gzi_[lo].first) != 18) { memset(dst + done, 0, n - done); return; }
      const size_t bsize = (size_t)(hdr[16] | (hdr[17] << 8)) + 1;   // BSIZE of the BC extra field
      std::vector<uint8_t> comp(bsize);
      if (pread(fd_, comp.data(), bsize, (off_t)gzi_[lo].first) != (ssize_t)bsize) { memset(dst + done, 0, n - done); return; }
      uint32_t isize; memcpy(&isize, comp.data() + bsize - 4, 4);
      block_.assign(isize, 0);
      z_stream zs{};
      if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("inflateInit2 failed");
      zs.next_in = comp.data() + 18; zs.avail_in = (uInt)(bsize - 18 - 8);
      zs.next_out = block_.data(); zs.avail_out = isize;
      const int rc = inflate(&zs, Z_FINISH);
      inflateEnd(&zs);
      if (rc != Z_STREAM_END) throw std::runtime_error("corrupt BGZF block in the reference genome");
      block_uoff_ = (int64_t)gzi_[lo].second;
      if (block_.empty() || u >= block_uoff_ + (int64_t)block_.size()) { memset(dst + done, 0, n - done); return; }   // past the end
    }
    const size_t k = std::min<size_t>(n - done, (size_t)(block_uoff_ + (int64_t)block_.size() - u));
    memcpy(dst + done, block_.data() + (u - block_uoff_), k);
    done += k;
  }
}

char IndexedFasta::base_at(const std::string& contig, int64_t pos) {
  if (contig == win_contig_ && pos >= win_pos_ && pos < win_pos_ + (int64_t)win_.size()) return win_[pos - win_pos_];
  auto it = index_.find(contig);
  if (it == index_.end()) throw std::runtime_error("contig " + contig + " is not in the reference genome");
  const Entry& e = it->second;
  if (pos < 0 || pos >= e.length) return 'N';
  const int64_t want = std::min<int64_t>(32768, e.length - pos);
  const int64_t start = e.offset + pos / e.line_bases * e.line_width + pos % e.line_bases;
  std::vector<char> raw((size_t)(want + want / e.line_bases * (e.line_width - e.line_bases) + 2 * e.line_width));
  read_raw(start, raw.data(), raw.size());
  win_.clear();
  for (char c : raw) { if ((int64_t)win_.size() >= want) break; if (isgraph((unsigned char)c)) win_.push_back(c); }
  win_contig_ = contig; win_pos_ = pos;
  return win_.empty() ? 'N' : win_[0];
}

}  // namespace genomicsdb_b200
